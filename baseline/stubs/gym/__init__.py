"""Minimal stand-in for the `gym` package (absent from this image), enough for the UNMODIFIED reference
`truck_trailer_sim/simv2.py` to import and run: `gym.Env` (simv2.py:20), `gym.spaces.Box` (simv2.py:79-91: only
.low / .high / .shape / .dtype are read) and `gym.error`.  Baseline harness only -- not product code."""
import numpy as np

from . import error, spaces  # noqa: F401


class Env:
    reward_range = (-float("inf"), float("inf"))
    metadata: dict = {}
