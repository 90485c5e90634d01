"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of ``oracle/tt_oracle.c`` (the CPU restatement of the
reference hot path).  See the header of tt_oracle.c for what it follows and how it is pinned.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtt_oracle.so")

NCOMP = 11
NFLAGS = 6
COMP_NAMES = ("total_reward", "distance_reward", "progress_reward", "heading_reward", "orientation_reward",
              "staged_success", "safety_penalty", "exploration_bonus", "final_success_bonus",
              "backward_penalty", "smoothness_penalty")
FLAG_NAMES = ("jackknife", "out_of_map", "max_steps_reached", "goal_reached", "goal_passed",
              "excessive_backward")
VIOLATION_NAMES = ("none", "jackknife", "jackknife_warning", "major_boundary", "minor_boundary",
                   "past_the_goal", "max_step", "excessive_backward")


# threshold margins reported by tto_step_m (positive = flag raised) and the termination flag each one decides
MARGIN_NAMES = ("jackknife", "out_of_map", "goal_passed", "excessive_backward", "goal_pos", "goal_ori", "max_steps")
NMARGINS = 7
# TT_F_* bit -> margins that decide it (goal_reached is the AND of two tests)
FLAG_MARGINS = {0: (0,), 1: (1,), 2: (6,), 3: (4, 5), 4: (2,), 5: (3,)}
# Documented epsilons of a "within-epsilon threshold crossing" (DESIGN.md section 3): the CUDA path carries angles in float64
# (observed error <= 2e-9 rad, growing with the unstable hitch dynamics) and positions in 2^-25 m fixed point with float32
# stage arithmetic (observed <= 4e-6 relative = 1.6e-4 m at map scale); max_steps is an integer test (no epsilon).
MARGIN_EPS = (1e-6, 2e-4, 2e-4, 2e-4, 2e-4, 2e-5, 0.0)


def explained_by_margin(flags_a, flags_b, margins):
    """True when every termination bit that differs between two flag bytes is a within-epsilon threshold crossing."""
    diff = int(flags_a) ^ int(flags_b)
    if diff == 0:
        return False
    for bit, ms in FLAG_MARGINS.items():
        if diff >> bit & 1 and not any(abs(margins[m]) < MARGIN_EPS[m] for m in ms):
            return False
    return True


class Cfg(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("L1", "L2", "v1x", "dt", "map_min", "map_max", "max_hitch",
                                           "steer_max", "pos_thr", "ori_thr", "step_len",
                                           "max_expected_distance")] + [("integrator", C.c_int)]


class Env(C.Structure):
    _fields_ = [("st", C.c_double * 6), ("startx", C.c_double), ("starty", C.c_double), ("startyaw", C.c_double),
                ("gx", C.c_double), ("gy", C.c_double), ("gyaw", C.c_double), ("steps", C.c_int), ("emax", C.c_int),
                ("has_rs", C.c_int), ("prev", C.c_double), ("cum", C.c_double), ("closest", C.c_double),
                ("hist", C.c_double * 5), ("first_steer", C.c_float), ("bt_steps", C.c_int), ("stage", C.c_int * 3)]


class Actor(C.Structure):
    _fields_ = [("in_dim", C.c_int), ("h1", C.c_int), ("h2", C.c_int)] + \
               [(n, C.c_void_p) for n in ("w1", "b1", "g1", "be1", "w2", "b2", "g2", "be2", "w3", "b3")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtt_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        assert L.tto_sizeof_env() == C.sizeof(Env), "oracle struct layout mismatch"
        L.tto_rollout_port.restype = C.c_int64
        L.tto_rng_normal.restype = C.c_float
        L.tto_rng_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.tto_rng_pose.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32] + [C.POINTER(C.c_double)] * 3
        _lib = L
    return _lib


def default_cfg(integrator: int = 0) -> Cfg:
    c = Cfg()
    lib().tto_default_cfg(C.byref(c))
    c.integrator = integrator
    return c


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def rk45(y0, delta, cfg=None):
    cfg = cfg or default_cfg()
    y0 = np.ascontiguousarray(y0, np.float64)
    out = np.zeros(6)
    ns, nf = C.c_int(), C.c_int()
    lib().tto_rk45(C.byref(cfg), _p(y0), C.c_double(delta), _p(out), C.byref(ns), C.byref(nf))
    return out, ns.value, nf.value


def dp5_fixed(y0, delta, cfg=None):
    cfg = cfg or default_cfg()
    y0 = np.ascontiguousarray(y0, np.float64)
    out = np.zeros(6)
    lib().tto_dp5_fixed(C.byref(cfg), _p(y0), C.c_double(delta), _p(out))
    return out


def observation(st, steer, goal, cfg=None):
    cfg = cfg or default_cfg()
    st = np.ascontiguousarray(st, np.float64)
    o = np.zeros(23, np.float32)
    lib().tto_obs(C.byref(cfg), _p(st), C.c_double(steer), C.c_double(goal[0]), C.c_double(goal[1]),
                  C.c_double(goal[2]), _p(o))
    return o


GOAL_DEFAULT = (0.0, -30.0, float(np.deg2rad(90)))


class OracleEnv:
    """One env with the reference's reset/step contract, backed by tto_*."""

    def __init__(self, cfg=None):
        self.cfg = cfg or default_cfg()
        self.e = Env()

    def reset_pose(self, sx, sy, syaw, goal=GOAL_DEFAULT):
        o = np.zeros(23, np.float32)
        lib().tto_reset_pose(C.byref(self.cfg), C.byref(self.e), *[C.c_double(float(v)) for v in (sx, sy, syaw, *goal)], _p(o))
        return o

    def set_state(self, st, start, goal=GOAL_DEFAULT):
        st = np.ascontiguousarray(st, np.float64)
        o = np.zeros(23, np.float32)
        lib().tto_set_state(C.byref(self.cfg), C.byref(self.e), _p(st),
                            *[C.c_double(float(v)) for v in (*start, *goal)], _p(o))
        return o

    @property
    def state(self):
        return np.array(self.e.st[:], np.float64)

    def step(self, action):
        o = np.zeros(23, np.float32)
        comps = np.zeros(NCOMP)
        viol, succ = C.c_int(), C.c_int()
        flags = (C.c_int * NFLAGS)()
        done = lib().tto_step(C.byref(self.cfg), C.byref(self.e), C.c_float(float(action)), _p(o), _p(comps),
                              C.byref(viol), flags, C.byref(succ))
        return o, comps, bool(done), viol.value, np.array(flags[:], np.uint8), bool(succ.value)

    def replay(self, actions):
        actions = np.ascontiguousarray(actions, np.float32).reshape(-1)
        T = len(actions)
        out = dict(state=np.zeros((T, 6)), obs=np.zeros((T, 23), np.float32), comps=np.zeros((T, NCOMP)),
                   viol=np.zeros(T, np.uint8), flags=np.zeros((T, NFLAGS), np.uint8), done=np.zeros(T, np.uint8),
                   success=np.zeros(T, np.uint8))
        n = lib().tto_replay(C.byref(self.cfg), C.byref(self.e), _p(actions), T, _p(out["state"]), _p(out["obs"]),
                             _p(out["comps"]), _p(out["viol"]), _p(out["flags"]), _p(out["done"]), _p(out["success"]))
        return {k: v[:n] for k, v in out.items()}

    def replay_m(self, actions):
        """``replay`` plus the per-step threshold margins [n, 7] (``MARGIN_NAMES``; positive = flag raised)."""
        actions = np.ascontiguousarray(actions, np.float32).reshape(-1)
        T = len(actions)
        out = dict(state=np.zeros((T, 6)), obs=np.zeros((T, 23), np.float32), comps=np.zeros((T, NCOMP)),
                   flags=np.zeros((T, NFLAGS), np.uint8), done=np.zeros(T, np.uint8), margins=np.zeros((T, NMARGINS)))
        n = lib().tto_replay_m(C.byref(self.cfg), C.byref(self.e), _p(actions), T, _p(out["state"]), _p(out["obs"]),
                               _p(out["comps"]), _p(out["flags"]), _p(out["done"]), _p(out["margins"]))
        return {k: v[:n] for k, v in out.items()}


ACTOR_KEYS = ("fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight",
              "bn2.bias", "mu.weight", "mu.bias")


class OracleActor:
    """networks.py:138-147 forward in float64 from a reference-layout state_dict of numpy arrays."""

    def __init__(self, sd):
        self.w = [np.ascontiguousarray(np.asarray(sd[k]), np.float32) for k in ACTOR_KEYS]
        self.a = Actor()
        self.a.h1, self.a.in_dim = self.w[0].shape
        self.a.h2 = self.w[4].shape[0]
        for name, arr in zip(("w1", "b1", "g1", "be1", "w2", "b2", "g2", "be2", "w3", "b3"), self.w):
            setattr(self.a, name, arr.ctypes.data)

    def forward(self, obs):
        obs = np.ascontiguousarray(obs, np.float32)
        B, ld = obs.shape
        out = np.zeros(B, np.float32)
        lib().tto_actor_forward(C.byref(self.a), _p(obs), ld, B, _p(out))
        return out


def philox(ctr, key):
    ctr = np.asarray(ctr, np.uint32); key = np.asarray(key, np.uint32); out = np.zeros(4, np.uint32)
    lib().tto_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def rng_pose(seed, gid, t):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib().tto_rng_pose(seed, gid, t, C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def rng_normal(seed, gid, t):
    return lib().tto_rng_normal(seed, gid, t)


def ou_step(x, action, reset_mask, seed, gid0, t):
    """In-place on x (float32[N]) and action (float32[N] or None)."""
    lib().tto_ou_step(_p(x), _p(action), _p(reset_mask), len(x), C.c_uint64(seed), C.c_uint32(gid0), C.c_uint32(t))


def rollout_replay(seed, gid0, a_raw, pose_t0=0x80000000, iter0=1, cfg=None, want_obs=True):
    """Open-loop oracle replay of a recorded rollout: ``a_raw`` float32 [T, n] = the ring's raw actions of envs
    gid0 .. gid0 + n - 1.  See tto_rollout_replay.  Returns arrays shaped [T, n, ...]."""
    cfg = cfg or default_cfg()
    a_raw = np.ascontiguousarray(a_raw, np.float32)
    T, n = a_raw.shape
    out = dict(rew=np.zeros((T, n)), done=np.zeros((T, n), np.uint8), flags=np.zeros((T, n), np.uint8),
               state=np.zeros((T, n, 6)), margins=np.zeros((T, n, NMARGINS)), ou=np.zeros((T, n), np.float32))
    if want_obs:
        out["s"] = np.zeros((T, n, 23), np.float32); out["s2"] = np.zeros((T, n, 23), np.float32)
    lib().tto_rollout_replay(C.byref(cfg), C.c_uint64(seed), C.c_uint32(gid0), C.c_uint32(pose_t0), C.c_uint32(iter0),
                             C.c_int64(n), C.c_int(T), _p(a_raw), _p(out.get("s")), _p(out.get("s2")), _p(out["rew"]),
                             _p(out["done"]), _p(out["flags"]), _p(out["state"]), _p(out["margins"]), _p(out["ou"]))
    return out


def replay_store(S, A, R, S2, D, cntr, s, a, r, s2, d):
    cap = S.shape[0]
    N = s.shape[0]
    lib().tto_replay_store(_p(S), _p(A), _p(R), _p(S2), _p(D), C.c_int64(cap), C.c_int64(cntr),
                           _p(s), s.strides[0] // 4, _p(a), _p(r), _p(s2), s2.strides[0] // 4, _p(d), C.c_int64(N))


class RolloutPort:
    """CPU-baseline port of the whole rollout iteration (bench.py only).  One slice per host thread."""

    def __init__(self, N, actor_sd, seed=27, threads=None, capacity=None, cfg=None):
        from concurrent.futures import ThreadPoolExecutor
        self.cfg = cfg or default_cfg()
        self.N, self.seed, self.t = N, seed, 0
        self.threads = threads or os.cpu_count() or 1
        self.pool = ThreadPoolExecutor(self.threads)
        self.envs = (Env * N)()
        self.obs = np.zeros((N, 23), np.float32)
        self.ou = np.zeros(N, np.float32)
        self.actor = OracleActor(actor_sd)
        cap = capacity or N
        self.cap, self.cntr = cap, 0
        self.S = np.zeros((cap, 23), np.float32); self.S2 = np.zeros((cap, 23), np.float32)
        self.A = np.zeros(cap, np.float32); self.R = np.zeros(cap, np.float32); self.D = np.zeros(cap, np.uint8)
        L = lib()
        for i in range(N):
            sx, sy, syaw = rng_pose(seed, i, 0xFFFFFFFF)
            L.tto_reset_pose(C.byref(self.cfg), C.byref(self.envs[i]), C.c_double(sx), C.c_double(sy),
                             C.c_double(syaw), C.c_double(0.0), C.c_double(-30.0), C.c_double(GOAL_DEFAULT[2]),
                             self.obs[i].ctypes.data_as(C.c_void_p))

    def step(self):
        L = lib()
        bounds = np.linspace(0, self.N, self.threads + 1).astype(np.int64)

        def run(k):
            return L.tto_rollout_port(C.byref(self.cfg), self.envs, _p(self.obs), _p(self.ou), C.byref(self.actor.a),
                                      C.c_int64(int(bounds[k])), C.c_int64(int(bounds[k + 1])), C.c_uint64(self.seed),
                                      C.c_uint32(0), C.c_uint32(self.t), _p(self.S), _p(self.A), _p(self.R),
                                      _p(self.S2), _p(self.D), C.c_int64(self.cap), C.c_int64(self.cntr), None)
        fin = sum(self.pool.map(run, range(self.threads)))
        self.t += 1
        self.cntr += self.N
        return fin
