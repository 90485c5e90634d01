"""Vectorised ``Truck_trailer_Env_2`` (truck_trailer_sim/simv2.py:20-545) on one B200.

``VecTruckTrailerEnv`` keeps the gym-style contract of the reference with a leading env dimension:
``reset(seed=None, options=None) -> (obs[N,23], {})`` and ``step(action) -> (obs, reward, done, info)``
(old-gym 4-tuple, ``done`` includes truncation, simv2.py:541-545).  ``Truck_trailer_Env_2`` is the N=1
numpy-facing wrapper that satisfies the reference driver loop (DDPG/trainv2.py:488-531) unchanged.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from ._lib import TT_NCOMP, TT_NSTATS, TT_OBS_DIM, COMP_NAMES, FLAG_NAMES, VIOLATION_NAMES, check, ptr, stream_ptr


class EnvConfig:
    """Environment constants; defaults = reference literals (simv2.py:23-101, :331-337)."""

    def __init__(self, **overrides):
        self.c = _lib.EnvCfg()
        check(_lib.load().tt_env_default_cfg(C.byref(self.c)))
        for k, v in overrides.items():
            if not hasattr(self.c, k):
                raise AttributeError(f"unknown env config field {k!r}")
            setattr(self.c, k, float(v))

    def __getattr__(self, name):
        return getattr(self.__dict__["c"], name)


class _Box:
    """Minimal stand-in for gym.spaces.Box (simv2.py:79-91): .low .high .shape .dtype"""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


class VecTruckTrailerEnv:
    """N independent truck-trailer environments stepped by one fused CUDA kernel.

    Parameters
    ----------
    num_envs : environments on this GPU.
    seed : Philox seed (reference default seed 27, trainv2.py:376).
    global_env_offset : global id of env 0 (multi-GPU sharding: rank * num_envs).
    ld_obs : row stride of observation buffers in floats (>= 23).
    emit_info : also produce the per-step reward components / violation / flags (the reference ``info``).
    auto_tick : a MASKED ``reset`` (the driver-side reset on done) also advances the Philox iteration counter, so that a
        hand-written loop ``choose_action -> step -> reset(options={'mask': done})`` is one iteration of the random
        streams (new start poses and new OU normals every pass).  With ``auto_tick=False`` the driver calls ``tick()``
        once per iteration itself.  (``RolloutEngine`` / ``tt_rollout_step`` tick inside the env kernel.)
    """

    metadata = {"render.modes": ["human", "rgb_array"]}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, num_envs: int, seed: int = 27, global_env_offset: int = 0, cfg: EnvConfig | None = None,
                 ld_obs: int = TT_OBS_DIM, emit_info: bool = False, device: str | torch.device | None = None,
                 auto_tick: bool = True):
        _lib.require_cuda()
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_envs = int(num_envs)
        self.cfg = cfg or EnvConfig()
        self.ld_obs = int(ld_obs)
        self.emit_info = emit_info
        self.auto_tick = bool(auto_tick)
        self.seed_value = int(seed)
        self.global_env_offset = int(global_env_offset)
        N = self.num_envs
        with torch.cuda.device(self.device):
            nbytes = self.L.tt_env_workspace_bytes(N)
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._ws_ptr = (self._ws.data_ptr() + 255) // 256 * 256
            h = C.c_void_p()
            check(self.L.tt_env_create(C.byref(h), C.byref(self.cfg.c), N, self.seed_value, self.global_env_offset,
                                       self._ws_ptr, nbytes))
            self._h = h
            self._obs = [torch.zeros(N, self.ld_obs, dtype=torch.float32, device=self.device) for _ in range(2)]
            self._per_env_l2 = False
            self._cur = 0
            self._reward = torch.zeros(N, dtype=torch.float32, device=self.device)
            self._done = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._done_bits = None
            self._scaled = torch.zeros(N, dtype=torch.float32, device=self.device)
            self._stats = torch.zeros(TT_NSTATS, dtype=torch.float64, device=self.device)
            self._info_bufs = None
        # reference attributes (simv2.py:23-101)
        c = self.cfg
        self.min_map_x = self.min_map_y = c.map_min
        self.max_map_x = self.max_map_y = c.map_max
        self.L1, self.v1x, self.dt = c.L1, c.v1x, c.dt
        self.position_threshold, self.orientation_threshold = c.pos_thr, c.ori_thr
        self.observation_dim = TT_OBS_DIM
        self.observation_space = _Box(-1.0, 1.0, (TT_OBS_DIM,), np.float32)
        self.action_space = _Box(-c.steer_max, c.steer_max, (1,), np.float32)
        self.single_observation_space, self.single_action_space = self.observation_space, self.action_space

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.L.tt_env_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def L2(self):
        """Trailer length: the configured scalar, or the per-env float64[N] tensor once ``set_l2`` was used."""
        return self.get_l2() if self._per_env_l2 else self.cfg.L2

    def set_l2(self, l2, idx=None):
        """``env.L2 = value`` per environment (heatmap.py:89 draws the trailer length per trial).  ``l2``: scalar or [n]
        float64; ``idx``: env indices (default 0..n-1).  Affects the dynamics from the next step on and the truck pose of
        the next ``reset`` / ``set_state``."""
        with torch.cuda.device(self.device):
            ix = None if idx is None else torch.as_tensor(idx, dtype=torch.int64, device=self.device).reshape(-1).contiguous()
            n = self.num_envs if ix is None else ix.numel()
            v = torch.as_tensor(l2, dtype=torch.float64, device=self.device).reshape(-1)
            if v.numel() == 1:
                v = v.expand(n)
            v = v.contiguous()
            if v.numel() != n or bool((v <= 0).any()):
                raise ValueError("set_l2: need one positive length per selected env")
            check(self.L.tt_env_set_l2(self._h, ptr(ix), n, v.data_ptr(), stream_ptr()))
            self._per_env_l2 = True

    def get_l2(self):
        with torch.cuda.device(self.device):
            out = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
            check(self.L.tt_env_get_l2(self._h, out.data_ptr(), stream_ptr()))
        return out

    def _obs_view(self, t):
        return t[:, :TT_OBS_DIM] if self.ld_obs != TT_OBS_DIM else t

    @property
    def obs_buffer(self):
        """The full [N, ld_obs] buffer holding the current observations."""
        return self._obs[self._cur]

    def _info_struct(self):
        if self._info_bufs is None:
            N, dev = self.num_envs, self.device
            self._info_bufs = SimpleNamespace(
                comps=torch.zeros(TT_NCOMP, N, dtype=torch.float32, device=dev),
                viol=torch.zeros(N, dtype=torch.uint8, device=dev), flags=torch.zeros(N, dtype=torch.uint8, device=dev),
                success=torch.zeros(N, dtype=torch.uint8, device=dev))
        b = self._info_bufs
        return _lib.StepInfo(b.comps.data_ptr(), b.viol.data_ptr(), b.flags.data_ptr(), b.success.data_ptr())

    # ------------------------------------------------------------------ reference API
    def reset(self, seed=None, options=None):
        """simv2.py:459-498.  ``options={'mask': done}`` resets only the envs whose mask is set (the
        driver-side reset on ``done``, trainv2.py:489) and returns the observation batch with those rows
        replaced; without a mask every env starts a new episode.  A masked reset ends one iteration of the Philox
        streams (see ``auto_tick``): call it once per loop pass, also when no env has finished."""
        with torch.cuda.device(self.device):
            if seed is not None:
                self.seed_value = int(seed)
                check(self.L.tt_env_seed(self._h, self.seed_value, stream_ptr()))
            mask = None if not options else options.get("mask")
            if mask is not None:
                mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            buf = self._obs[self._cur]
            check(self.L.tt_env_reset(self._h, ptr(mask), buf.data_ptr(), self.ld_obs, stream_ptr()))
            if mask is not None and self.auto_tick:
                check(self.L.tt_env_tick(self._h, 1, stream_ptr()))
        return self._obs_view(buf), {}

    def step(self, action, emit_info=None):
        """simv2.py:499-545 for every env.  ``action``: scaled steering, shape [N] or [N,1] (CUDA float32).
        Returns ``(obs[N,23], reward[N], done[N] bool, info)``; the returned observation of a finished env
        is its TERMINAL observation (what trainv2.py:525 stores); call ``reset(options={'mask': done})``
        afterwards like the reference driver does."""
        emit = self.emit_info if emit_info is None else emit_info
        with torch.cuda.device(self.device):
            a = torch.as_tensor(action, dtype=torch.float32, device=self.device).reshape(-1).contiguous()
            if a.numel() != self.num_envs:
                raise ValueError(f"action has {a.numel()} elements, expected {self.num_envs}")
            self._cur ^= 1
            buf = self._obs[self._cur]
            info_struct = self._info_struct() if emit else None
            check(self.L.tt_env_step(self._h, a.data_ptr(), buf.data_ptr(), self.ld_obs, self._reward.data_ptr(),
                                     self._done.data_ptr(), C.byref(info_struct) if emit else None, stream_ptr()))
        info = self._make_info() if emit else {}
        return self._obs_view(buf), self._reward, self._done.bool(), info

    def set_done_bits(self, bits):
        """Optional bit-packed copy of ``done`` (``tt_env_set_done_bits``): while set, every single-step launch also writes
        bit i % 32 of word i // 32 = done of env i into ``bits`` (int32 / uint32 [ceil(N / 32)], CUDA) -- 1 bit instead of 1 byte
        per env for a host that reads the flags back every step.  ``None`` switches it off.  ``unpack_done_bits`` decodes."""
        if bits is not None:
            if bits.device != self._done.device or bits.element_size() != 4 or bits.numel() < (self.num_envs + 31) // 32 or not bits.is_contiguous():
                raise ValueError("done bits: a contiguous 32-bit integer CUDA tensor with ceil(num_envs / 32) elements")
        self._done_bits = bits                          # keeps the buffer alive while the env points at it
        check(self.L.tt_env_set_done_bits(self._h, None if bits is None else bits.data_ptr()))

    @staticmethod
    def unpack_done_bits(bits, n):
        """bool [n] from the bit-packed words (numpy array or CPU / CUDA tensor)."""
        import numpy as np
        a = bits.detach().cpu().numpy() if hasattr(bits, "detach") else np.asarray(bits)
        return np.unpackbits(a.view(np.uint8), bitorder="little")[:n].astype(bool)

    def tick(self, by: int = 1):
        """Advance the Philox iteration counter (once per step+reset iteration when driving by hand)."""
        with torch.cuda.device(self.device):
            check(self.L.tt_env_tick(self._h, by, stream_ptr()))

    def step_k(self, actions, auto_reset=False, want_obs=False, want_reward=True, want_done=True, want_info=False):
        """K consecutive steps in one launch (state in registers); ``actions`` [K, N] scaled steering."""
        with torch.cuda.device(self.device):
            a = torch.as_tensor(actions, dtype=torch.float32, device=self.device).contiguous()
            K, N = a.shape
            assert N == self.num_envs
            dev = self.device
            obs = torch.empty(K, N, self.ld_obs, dtype=torch.float32, device=dev) if want_obs else None
            rew = torch.empty(K, N, dtype=torch.float32, device=dev) if want_reward else None
            done = torch.empty(K, N, dtype=torch.uint8, device=dev) if want_done else None
            info = None
            if want_info:
                info = SimpleNamespace(comps=torch.empty(K, TT_NCOMP, N, dtype=torch.float32, device=dev),
                                       viol=torch.empty(K, N, dtype=torch.uint8, device=dev),
                                       flags=torch.empty(K, N, dtype=torch.uint8, device=dev),
                                       success=torch.empty(K, N, dtype=torch.uint8, device=dev))
                st = _lib.StepInfo(info.comps.data_ptr(), info.viol.data_ptr(), info.flags.data_ptr(), info.success.data_ptr())
            check(self.L.tt_env_step_k(self._h, a.data_ptr(), K, int(auto_reset), ptr(obs), self.ld_obs, ptr(rew), ptr(done),
                                       C.byref(st) if want_info else None, stream_ptr()))
        return obs, rew, done, info

    def _make_info(self):
        b = self._info_bufs
        info = {name: b.comps[i] for i, name in enumerate(COMP_NAMES)}
        info["total_reward"] = self._reward
        info["violation_type"] = b.viol          # codes, see VIOLATION_NAMES (reward_functionv1.py:378-419 order)
        info["success"] = b.success.bool()
        info["termination_flags"] = b.flags      # bit i = FLAG_NAMES[i]
        return info

    # ------------------------------------------------------------------ state injection / readback
    def set_state(self, state, start, goal=None, idx=None):
        """Inject states (test.py:96-115, heatmap.py:119): state [n,6] = psi1,psi2,x1,y1,x2,y2; start [n,3];
        goal [n,3] (default: reference goal).  Starts fresh episodes; returns the observation batch."""
        with torch.cuda.device(self.device):
            dev = self.device
            st = torch.as_tensor(np.asarray(state, np.float64) if not torch.is_tensor(state) else state, dtype=torch.float64, device=dev).reshape(-1, 6).contiguous()
            n = st.shape[0]
            sp = torch.as_tensor(np.asarray(start, np.float64) if not torch.is_tensor(start) else start, dtype=torch.float64, device=dev).reshape(-1, 3).contiguous()
            if goal is None:
                goal = np.tile([self.cfg.goal_x, self.cfg.goal_y, self.cfg.goal_yaw], (n, 1))
            gl = torch.as_tensor(np.asarray(goal, np.float64) if not torch.is_tensor(goal) else goal, dtype=torch.float64, device=dev).reshape(-1, 3).contiguous()
            ix = None if idx is None else torch.as_tensor(idx, dtype=torch.int64, device=dev).contiguous()
            buf = self._obs[self._cur]
            check(self.L.tt_env_set_state(self._h, ptr(ix), n, st.data_ptr(), sp.data_ptr(), gl.data_ptr(), buf.data_ptr(),
                                          self.ld_obs, stream_ptr()))
        return self._obs_view(buf)

    def get_state(self):
        """Returns dict(state[N,6] f64, start[N,3], goal[N,3], episode_steps[N], max_episode_steps[N])."""
        with torch.cuda.device(self.device):
            N, dev = self.num_envs, self.device
            st = torch.empty(N, 6, dtype=torch.float64, device=dev)
            sp = torch.empty(N, 3, dtype=torch.float64, device=dev)
            gl = torch.empty(N, 3, dtype=torch.float64, device=dev)
            steps = torch.empty(N, dtype=torch.int32, device=dev)
            ms = torch.empty(N, dtype=torch.int32, device=dev)
            check(self.L.tt_env_get_state(self._h, st.data_ptr(), sp.data_ptr(), gl.data_ptr(), steps.data_ptr(), ms.data_ptr(),
                                          stream_ptr()))
        return dict(state=st, start=sp, goal=gl, episode_steps=steps, max_episode_steps=ms)

    def get_reward_state(self):
        """The reward function's persistent per-episode state (reward_functionv1.py:99-109) as float32 [N] tensors:
        closest_distance_to_goal, cumulative_backward_movement, previous_steering (frozen first steering), episode_return."""
        with torch.cuda.device(self.device):
            out = [torch.empty(self.num_envs, dtype=torch.float32, device=self.device) for _ in range(4)]
            check(self.L.tt_env_get_reward_state(self._h, *[t.data_ptr() for t in out], stream_ptr()))
        return dict(zip(("closest_distance_to_goal", "cumulative_backward_movement", "previous_steering", "episode_return"), out))

    @property
    def state(self):
        return self.get_state()["state"]

    def read_stats(self, clear=True):
        """Device-side rollout statistics since the last read -> dict of python floats (one D2H copy)."""
        with torch.cuda.device(self.device):
            check(self.L.tt_env_stats_read(self._h, self._stats.data_ptr(), int(clear), stream_ptr()))
            v = self._stats.cpu().tolist()
        return dict(zip(_lib.STAT_NAMES, v))

    def stats_tensor(self, clear=True):
        """Same statistics as a CUDA float64[16] tensor (for an NCCL all-reduce without host round trip)."""
        with torch.cuda.device(self.device):
            check(self.L.tt_env_stats_read(self._h, self._stats.data_ptr(), int(clear), stream_ptr()))
        return self._stats

    def compute_max_steps(self):
        return self.get_state()["max_episode_steps"]

    def render(self, mode="human"):      # rendering is out of scope of the hot path (simv2.py:547-605)
        return None

    def close(self):
        return None


class Truck_trailer_Env_2:
    """N=1 numpy-facing environment with the exact caller contract of the reference class of the same name
    (simv2.py:20; used by DDPG/trainv2.py:488-531, test.py, heatmap.py): numpy observation (23,), python
    float reward, bool done, dict info with the reference keys; attributes ``state``, ``startx`` ... are
    readable and assignable.  Backed by the same CUDA kernels as the vector env."""

    metadata = VecTruckTrailerEnv.metadata
    reward_range = VecTruckTrailerEnv.reward_range

    def __init__(self, seed: int = 27, device=None):
        self.vec = VecTruckTrailerEnv(1, seed=seed, emit_info=True, device=device, auto_tick=False)
        v = self.vec
        self.observation_space, self.action_space = v.observation_space, v.action_space
        self.observation_dim = TT_OBS_DIM
        self.position_threshold, self.orientation_threshold = v.position_threshold, v.orientation_threshold
        self.min_map_x, self.min_map_y, self.max_map_x, self.max_map_y = v.min_map_x, v.min_map_y, v.max_map_x, v.max_map_y
        self.L1, self._L2, self.v1x, self.dt = v.L1, float(v.cfg.L2), v.v1x, v.dt
        self.path_x = self.path_y = self.path_yaw = []       # read by DDPG/train.py:362-364 (simv1 legacy)
        self.jackknife = self.out_of_map = self.max_steps_reached = self.goal_passed = self.goal_reached = False
        self._episode = 0
        self._init_pack()
        self._sync_pose()

    # Everything one ``step`` hands back (observation, reward, done, the 13 reward components, flags, the state and the reward
    # function's persistent state) lives in ONE device block and comes back in ONE pinned copy: the reference contract wants ~20
    # python scalars per step, and reading them one by one was 22 host synchronisations (386 us per call; now one).
    _PACK = (("obs", np.float32, TT_OBS_DIM), ("reward", np.float32, 1), ("comps", np.float32, TT_NCOMP), ("state", np.float64, 6),
             ("start", np.float64, 3), ("goal", np.float64, 3), ("steps", np.int32, 1), ("max_steps", np.int32, 1),
             ("rs", np.float32, 4), ("done", np.uint8, 1), ("viol", np.uint8, 1), ("flags", np.uint8, 1), ("success", np.uint8, 1))

    def _init_pack(self):
        dev = self.vec.device
        off, self._poff = 0, {}
        for name, dt, n in self._PACK:
            self._poff[name] = off
            off += (np.dtype(dt).itemsize * n + 15) // 16 * 16
        with torch.cuda.device(dev):
            self._pack_dev = torch.zeros(off, dtype=torch.uint8, device=dev)
            self._pack_host = torch.zeros(off, dtype=torch.uint8).pin_memory()
            self._a_host = torch.zeros(1, dtype=torch.float32).pin_memory()
            self._a_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        base, host = self._pack_dev.data_ptr(), self._pack_host.numpy()
        self._pptr = {name: base + o for name, o in self._poff.items()}
        self._pview = {name: host[self._poff[name]:self._poff[name] + np.dtype(dt).itemsize * n].view(dt) for name, dt, n in self._PACK}
        p = self._pptr
        self._pinfo = _lib.StepInfo(p["comps"], p["viol"], p["flags"], p["success"])

    def _sync_pose(self):
        s = self.vec.get_state()
        self._state = s["state"][0].cpu().numpy()
        sp, gl = s["start"][0].cpu().numpy(), s["goal"][0].cpu().numpy()
        self.startx, self.starty, self.startyaw = float(sp[0]), float(sp[1]), float(sp[2])
        self.goalx, self.goaly, self.goalyaw = float(gl[0]), float(gl[1]), float(gl[2])
        self.episode_steps = int(s["episode_steps"][0])
        self.max_episode_steps = int(s["max_episode_steps"][0])

    @property
    def L2(self):
        return self._L2

    @L2.setter
    def L2(self, value):
        # heatmap.py:89: `env.L2 = np.random.uniform(5, 7)` before the start state of the trial is assigned
        self._L2 = float(value)
        self.vec.set_l2(self._L2)

    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):
        # test.py:111 / heatmap.py:119 / episode_replay_collectorv2.py:258 assign env.state after setting
        # startx/starty/startyaw/goal*: inject and start a fresh episode from it
        self._state = np.asarray(value, np.float64).copy()
        self.vec.set_state(self._state[None], [[self.startx, self.starty, self.startyaw]], [[self.goalx, self.goaly, self.goalyaw]])
        self._sync_pose()

    def compute_max_steps(self):
        return int(math.hypot(self.goalx - self.startx, self.goaly - self.starty) / 0.40096) + 75

    def reset(self, seed=None, options=None):
        if seed is not None:
            obs, _ = self.vec.reset(seed=seed)
        else:
            obs, _ = self.vec.reset()
        self._sync_pose()
        return obs[0].cpu().numpy().copy(), {}

    def step(self, action):
        v, L, p, out_v = self.vec, self.vec.L, self._pptr, self._pview
        self._a_host[0] = float(np.asarray(action, np.float32).reshape(-1)[0])
        with torch.cuda.device(v.device):
            st = stream_ptr()
            self._a_dev.copy_(self._a_host, non_blocking=True)
            check(L.tt_env_step(v._h, self._a_dev.data_ptr(), p["obs"], TT_OBS_DIM, p["reward"], p["done"], C.byref(self._pinfo), st))
            check(L.tt_env_tick(v._h, 1, st))
            check(L.tt_env_get_state(v._h, p["state"], p["start"], p["goal"], p["steps"], p["max_steps"], st))
            rs = p["rs"]
            check(L.tt_env_get_reward_state(v._h, rs, rs + 4, rs + 8, rs + 12, st))
            self._pack_host.copy_(self._pack_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        flags = int(out_v["flags"][0])
        (self.jackknife, self.out_of_map, self.max_steps_reached, self.goal_reached, self.goal_passed, _exb) = \
            [bool(flags >> i & 1) for i in range(6)]
        comps = out_v["comps"]
        out = {k: float(comps[i]) for i, k in enumerate(COMP_NAMES)}
        out["total_reward"] = float(out_v["reward"][0])
        out["violation_type"] = VIOLATION_NAMES[int(out_v["viol"][0])]
        out["success"] = bool(out_v["success"][0])
        self._state = out_v["state"].copy()
        self.episode_steps = int(out_v["steps"][0])
        # reward_functionv1.py:250-283 `penalty_info` (the 14th key of the reference's info dict)
        cum = float(out_v["rs"][1])
        budget = 5.0 * min(1.0, self.episode_steps / 50)
        out["backward_movement_info"] = {"cumulative_backward": cum, "movement_budget": budget,
                                         "excess_movement": max(0.0, cum - budget), "penalty": out["backward_penalty"]}
        return out_v["obs"].copy(), out["total_reward"], bool(out_v["done"][0]), out

    def compute_observation(self, state, steering_angle):
        """simv2.py:103-181 for an arbitrary state (heatmap.py:122 calls it on the state it has just assigned): computed by
        the state-injection kernel of a scratch env with the current start / goal / L2; the steering angle only enters the
        observation as sin/cos at indices 10, 11."""
        if getattr(self, "_scratch", None) is None:
            self._scratch = VecTruckTrailerEnv(1, seed=0, device=self.vec.device)
        sc = self._scratch
        sc.set_l2(self._L2)
        obs = sc.set_state(np.asarray(state, np.float64)[None], [[self.startx, self.starty, self.startyaw]],
                           [[self.goalx, self.goaly, self.goalyaw]])[0].cpu().numpy().copy()
        obs[10], obs[11] = np.float32(math.sin(steering_angle)), np.float32(math.cos(steering_angle))
        return obs

    def render(self, mode="human"):
        return None

    def close(self):
        return None
