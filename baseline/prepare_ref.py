"""Places the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, travels to the GPU box with the
snapshot like the built .so files) so that bench.py can time the reference's own Python loop on the GPU box's host cores.

The reference (pain7576/ddpg-trucktrailer) is a set of plain scripts without setup.py / pyproject, so there is nothing for
`pip install --target baseline/_ref` to install; this is a byte-for-byte copy of the seven files the loop
DDPG/trainv2.py:488-531 needs.  Run by __graft_entry__.build() wherever /root/reference is mounted."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TT_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["truck_trailer_sim/__init__.py", "truck_trailer_sim/simv2.py", "truck_trailer_sim/reward_functionv1.py",
         "DDPG/DDPG_agent.py", "DDPG/networks.py", "DDPG/noise.py", "DDPG/replay_buffer.py"]


def available() -> bool:
    return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)


def prepare() -> str | None:
    if not os.path.isfile(os.path.join(REF, FILES[1])):
        return DST if available() else None
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "sha256": manifest}, fh, indent=1)
    return DST


if __name__ == "__main__":
    print(prepare())
