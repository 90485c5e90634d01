"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads, exports every symbol
include/tt_b200.h declares, validates arguments, and FAILS LOUDLY without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libtt_b200.so does not export {n}"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert L.tt_abi_version() == 1


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", tt.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_default_cfg_matches_reference_literals():
    c = tt.EnvConfig()
    assert (c.L1, c.L2, c.v1x, c.dt) == (5.0, 7.0, -5.012, 0.08)           # simv2.py:34-40
    assert (c.map_min, c.map_max, c.pos_thr, c.step_len) == (-40.0, 40.0, 0.5, 0.40096)
    assert abs(c.steer_max - 0.7853981633974483) < 1e-16 and abs(c.ori_thr - 0.2617993877991494) < 1e-16
    assert (c.goal_x, c.goal_y) == (0.0, -30.0) and abs(c.goal_yaw - 1.5707963267948966) < 1e-16


def test_argument_validation_without_compute():
    L = _lib.load()
    assert L.tt_env_workspace_bytes(0) == 0 and L.tt_env_workspace_bytes(1 << 20) > (1 << 20) * 100
    h = C.c_void_p()
    cfg = tt.EnvConfig()
    assert L.tt_env_create(C.byref(h), C.byref(cfg.c), 16, 1, 0, None, 0) == -1          # NULL workspace
    assert b"NULL" in L.tt_last_error()
    assert L.tt_env_create(C.byref(h), C.byref(cfg.c), 16, 1, 0, 256, 16) == -3          # workspace too small
    assert L.tt_env_step(None, None, None, 23, None, None, None, None) == -1
    assert L.tt_actor_forward(None, None, 23, 1, None, 0, None) == -1
    assert L.tt_replay_store(None, None, None, None, None, 10, 0, None, 23, None, None, None, 23, None, 1, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_cuda():
    for ctor in (lambda: tt.VecTruckTrailerEnv(4), lambda: tt.DeviceReplayBuffer(16),
                 lambda: tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=4)):
        with pytest.raises(tt.TTError):
            ctor()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no file of the product package may import, load or link it."""
    pkg = os.path.join(ROOT, "ddpg-trucktrailer_b200")
    bad = re.compile(r"^\s*(import|from)\s+oracle\b|libtt_oracle|oracle\.oracle|#include.*oracle", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not bad.search(open(os.path.join(dp, f)).read()), f
