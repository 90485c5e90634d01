"""Numerical behaviour of the kernel arithmetic (csrc/tt_env_math.cuh) measured on the CPU: the header is
compiled for the host (tests/host_math/host_step.cpp, g++) and replayed against the golden fixtures from the
untouched reference.  This validates the mixed float64/float32 integration scheme before any GPU time is
spent; the CUDA build of the same header is checked on the B200 by tests/test_gpu_env.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "host_step.cpp")
LIB = os.path.join(HERE, "host_math", "libhost_math.so")


@pytest.fixture(scope="module")
def hm():
    deps = [SRC, os.path.join(HERE, "..", "ddpg-trucktrailer_b200", "csrc", "tt_env_math.cuh"),
            os.path.join(HERE, "..", "ddpg-trucktrailer_b200", "csrc", "tt_consts.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-mavx2", "-mfma", "-ffp-contract=fast", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    L = C.CDLL(LIB)
    L.hm_rng_normal.restype = C.c_float
    L.hm_rng_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.hm_rng_pose.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
    L.hm_reset_pose.argtypes = [C.c_double] * 3 + [C.c_void_p] * 2
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def replay(L, state0, start, goal, actions):
    T = len(actions)
    out = dict(state=np.zeros((T, 6)), obs=np.zeros((T, 23), np.float32), comps=np.zeros((T, 11), np.float32),
               viol=np.zeros(T, np.uint8), flags=np.zeros(T, np.uint8), done=np.zeros(T, np.uint8), success=np.zeros(T, np.uint8))
    obs0 = np.zeros(23, np.float32)
    n = L.hm_replay(_p(np.ascontiguousarray(state0, np.float64)), _p(np.ascontiguousarray(start, np.float64)),
                    _p(np.ascontiguousarray(goal, np.float64)), _p(np.ascontiguousarray(actions, np.float32)), T,
                    _p(out["state"]), _p(out["obs"]), _p(out["comps"]), _p(out["viol"]), _p(out["flags"]), _p(out["done"]),
                    _p(out["success"]), _p(obs0))
    o = {k: v[:n] for k, v in out.items()}
    o["obs0"] = obs0
    return o


def test_trig_kernels(hm):
    xs = np.random.default_rng(0).uniform(-40, 40, 5000)
    for x in xs:
        s, c = C.c_double(), C.c_double()
        hm.hm_sincos_f64(C.c_double(x), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(x)) < 3e-16 and abs(c.value - np.cos(x)) < 3e-16
        sf, cf = C.c_float(), C.c_float()
        hm.hm_sincos_f32(C.c_double(x), C.byref(sf), C.byref(cf))
        assert abs(sf.value - np.sin(x)) < 1.5e-7 and abs(cf.value - np.cos(x)) < 1.5e-7


def test_recorded_episode_10579(hm, golden_dir):
    g = np.load(os.path.join(golden_dir, "episode_10579.npz"))
    o = replay(hm, g["states"][0], g["start"], g["goal"], g["actions"])
    assert len(o["done"]) == 193 and o["done"][-1] == 1 and o["success"][-1] == 1 and not o["done"][:-1].any()
    ref = g["states"][1:]
    rel = np.abs(o["state"] - ref) / np.maximum(np.abs(ref), 1.0)
    assert rel.max() < 1e-5, rel.max()                        # north_star bar: 1e-4 relative
    assert np.abs(o["state"][:, :2] - ref[:, :2]).max() < 1e-7   # the unstable angle dynamics are float64
    assert np.abs(o["comps"] - g["comps"]).max() < 5e-4
    assert abs(o["comps"][:, 0].astype(np.float64).sum() - 4792.9998) < 2e-2
    assert np.array_equal(o["viol"], g["viol"])


def test_reference_rollouts(hm, golden_dir):
    R = np.load(os.path.join(golden_dir, "ref_rollouts.npz"))
    bits = np.array([1, 2, 4, 8, 16, 32])
    for i in range(len(R["length"])):
        n = int(R["length"][i])
        o = replay(hm, R["state0"][i], R["start"][i], R["goal"][i], R["actions"][i, :n])
        tag = (i, str(R["tag"][i]))
        assert len(o["done"]) == n and np.array_equal(o["done"], R["done"][i, :n]), tag
        ref = R["state"][i, :n]
        assert (np.abs(o["state"] - ref) / np.maximum(np.abs(ref), 1.0)).max() < 1e-5, tag
        assert np.abs(o["obs"] - R["obs"][i, :n]).max() < 5e-6, tag
        assert np.abs(o["obs0"] - R["obs0"][i]).max() < 5e-7, tag
        refc = R["comps"][i, :n]
        assert (np.abs(o["comps"] - refc) / np.maximum(np.abs(refc), 1.0)).max() < 1e-4, tag
        assert np.array_equal(o["flags"], (R["flags"][i, :n] * bits).sum(1).astype(np.uint8)), tag
        assert np.array_equal(o["viol"], R["viol"][i, :n]) and np.array_equal(o["success"], R["success"][i, :n]), tag
        assert hm.hm_max_steps(_p(np.ascontiguousarray(R["start"][i])), _p(np.ascontiguousarray(R["goal"][i]))) == R["max_steps"][i]


def test_reset_and_rng_match_oracle_spec(hm, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_resets.npz"))
    for pose, st, obs in zip(g["pose"], g["state"], g["obs"]):
        s = np.zeros(6); o = np.zeros(23, np.float32)
        hm.hm_reset_pose(*[float(v) for v in pose], _p(s), _p(o))
        assert np.array_equal(s[:2].astype(np.float32), st[:2]) and np.abs(s[2:] - st[2:]).max() <= 2.0 ** -26   # positions: 2^-25 m fixed point
        assert np.abs(o - obs).max() < 5e-7
    for gid, t in ((0, 0), (5, 17), (2 ** 31 + 3, 2 ** 32 - 1)):
        out = np.zeros(3)
        hm.hm_rng_pose(27, gid, t, _p(out))
        assert tuple(out) == orc.rng_pose(27, gid, t)           # bit-exact: integer Philox + exact float64 fma
        assert abs(hm.hm_rng_normal(27, gid, t) - orc.rng_normal(27, gid, t)) < 1e-5
