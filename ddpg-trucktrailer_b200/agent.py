"""Vectorised DDPG agent rollout side: ``Agent.choose_action / remember / noise.reset`` of
DDPG/DDPG_agent.py:9-52 for N observations at once.

The actor forward (DDPG/networks.py:138-147), the per-env Ornstein-Uhlenbeck noise (DDPG/noise.py) and the
replay store run as CUDA kernels behind the C ABI.  ``learn()`` (DDPG_agent.py:72-106)
is ``tt_learn_step``: hand-written kernels on the device ring (learner.py, csrc/tt_learn.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import PRECISIONS, TT_OBS_DIM, check, ptr, stream_ptr
from .replay import DeviceReplayBuffer

ACTOR_KEYS = ("fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight", "bn2.bias",
              "mu.weight", "mu.bias")


def init_actor_state_dict(input_dims=TT_OBS_DIM, fc1_dims=400, fc2_dims=300, n_actions=1, seed=None):
    """Initial actor parameters with the reference's distributions and draw order (networks.py:110-131):
    fc1/fc2 U(+-1/sqrt(out_features)) (the reference uses weight.size()[0]), mu U(+-0.003), LayerNorm (1, 0).
    Built on the CPU generator so that ``torch.manual_seed(s)`` gives the same tensors as the reference."""
    import torch.nn as nn
    if seed is not None:
        torch.manual_seed(seed)
    fc1 = nn.Linear(input_dims, fc1_dims)
    fc2 = nn.Linear(fc1_dims, fc2_dims)
    bn1, bn2 = nn.LayerNorm(fc1_dims), nn.LayerNorm(fc2_dims)
    mu = nn.Linear(fc2_dims, n_actions)
    f2 = 1.0 / np.sqrt(fc2.weight.data.size()[0])
    fc2.weight.data.uniform_(-f2, f2); fc2.bias.data.uniform_(-f2, f2)
    f1 = 1.0 / np.sqrt(fc1.weight.data.size()[0])
    fc1.weight.data.uniform_(-f1, f1); fc1.bias.data.uniform_(-f1, f1)
    mu.weight.data.uniform_(-0.003, 0.003); mu.bias.data.uniform_(-0.003, 0.003)
    mods = {"fc1": fc1, "bn1": bn1, "fc2": fc2, "bn2": bn2, "mu": mu}
    return {k: getattr(mods[k.split(".")[0]], k.split(".")[1]).data.clone() for k in ACTOR_KEYS}


class CudaActor:
    """Packed device copy of an ``ActorNetwork`` state_dict + the forward kernels."""

    def __init__(self, input_dims=TT_OBS_DIM, fc1_dims=400, fc2_dims=300, device=None):
        _lib.require_cuda()
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.dims = (int(input_dims), int(fc1_dims), int(fc2_dims))
        with torch.cuda.device(self.device):
            nbytes = self.L.tt_actor_workspace_bytes(*self.dims)
            self._ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            h = C.c_void_p()
            check(self.L.tt_actor_create(C.byref(h), *self.dims, (self._ws.data_ptr() + 255) // 256 * 256, nbytes))
            self._h = h
        self._sd = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.L.tt_actor_destroy(h)
            except Exception:
                pass
            self._h = None

    def load_state_dict(self, sd):
        """``sd``: reference ActorNetwork state_dict (torch tensors or numpy arrays, any device)."""
        i, h1, h2 = self.dims
        shapes = {"fc1.weight": (h1, i), "fc1.bias": (h1,), "bn1.weight": (h1,), "bn1.bias": (h1,), "fc2.weight": (h2, h1),
                  "fc2.bias": (h2,), "bn2.weight": (h2,), "bn2.bias": (h2,), "mu.weight": (1, h2), "mu.bias": (1,)}
        dev = {}
        for k in ACTOR_KEYS:
            t = torch.as_tensor(sd[k]).detach().to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != shapes[k]:
                raise ValueError(f"{k}: expected shape {shapes[k]}, got {tuple(t.shape)}")
            dev[k] = t
        with torch.cuda.device(self.device):
            check(self.L.tt_actor_load(self._h, *[dev[k].data_ptr() for k in ACTOR_KEYS], stream_ptr()))
        self._sd = dev

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self._sd.items()}

    def flat_size(self):
        i, h1, h2 = self.dims
        return h1 * i + 3 * h1 + h2 * h1 + 3 * h2 + h2 + 1

    def load_flat(self, flat):
        """Load from ONE flat float32 CUDA vector holding the ten tensors in ``ACTOR_KEYS`` order (the layout of the learner's
        actor parameters and of the broadcast message, include/tt_b200.h): no copies, the re-pack kernels read it in place."""
        i, h1, h2 = self.dims
        if flat.dtype != torch.float32 or flat.device != self.device or flat.numel() != self.flat_size() or not flat.is_contiguous():
            raise ValueError("load_flat: need a contiguous float32 vector of flat_size() elements on the actor's device")
        shapes = ((h1, i), (h1,), (h1,), (h1,), (h2, h1), (h2,), (h2,), (h2,), (1, h2), (1,))
        views, o = {}, 0
        for k, shp in zip(ACTOR_KEYS, shapes):
            n = int(np.prod(shp))
            views[k] = flat[o:o + n].view(shp)
            o += n
        with torch.cuda.device(self.device):
            check(self.L.tt_actor_load(self._h, *[views[k].data_ptr() for k in ACTOR_KEYS], stream_ptr()))
        self._sd = views

    def auto_precision(self, n):
        """What precision="auto" resolves to for a batch of n rows: "f16" (tcgen05 tiles) once the batch is a real dense
        contraction, "fp32" (warp-level FMA) below."""
        return {0: "fp32", 2: "f16"}[self.L.tt_actor_auto_precision(self._h, int(n))]

    def forward(self, obs, out=None, precision="fp32", allow_out_of_bar=False):
        """tanh(mu(relu(LN(fc2(relu(LN(fc1(obs)))))))) -> [n] float32 (networks.py:138-147).
        precision: "fp32" (CUDA cores, <=1e-5), "f16" (tcgen05, fp16 operands + exact first layer, <=1e-3), "auto" (one of
        those two by batch size); with ``allow_out_of_bar=True`` also "f16_plain" / "bf16" (tcgen05, plain 16-bit operands:
        faster, but above the 1e-3 bar on strongly amplified weights)."""
        with torch.cuda.device(self.device):
            if obs.dim() == 1:
                obs = obs.reshape(1, -1)
            if obs.dtype != torch.float32 or obs.stride(1) != 1 or obs.device != self.device:
                obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
            n = obs.shape[0]
            if out is None:
                out = torch.empty(n, dtype=torch.float32, device=self.device)
            _, prec = _lib.resolve_precision(precision, allow_out_of_bar)
            check(self.L.tt_actor_forward(self._h, obs.data_ptr(), obs.stride(0), n, out.data_ptr(), prec, stream_ptr()))
        return out


class OUNoiseState:
    """Per-env ``OUActionNoise`` (DDPG/noise.py:4-20): theta 0.2, sigma 0.15, dt 1e-2, mu 0; Philox normals
    keyed by (seed, global env id, iteration)."""

    def __init__(self, num_envs, device, seed=27, global_env_offset=0):
        self.x_prev = torch.zeros(num_envs, dtype=torch.float32, device=device)
        self.seed, self.global_env_offset = int(seed), int(global_env_offset)
        self._iter = torch.zeros(1, dtype=torch.int32, device=device)     # used when no env supplies the counter
        self.iter_ptr = self._iter.data_ptr()
        self._own_iter = True

    def bind_env(self, env):
        """Share the env's device-side iteration counter and seed (the reference's noise stream is a function
        of the episode seed, DDPG/trainv2.py:489 + noise.py:14)."""
        self.iter_ptr = env.L.tt_env_iter_ptr(env._h)
        self.seed, self.global_env_offset = env.seed_value, env.global_env_offset
        self._own_iter = False

    def reset(self, mask=None):
        """noise.py:19-20 (``agent.noise.reset()``, trainv2.py:492); ``mask`` restricts it to finished envs."""
        if mask is None:
            self.x_prev.zero_()
        else:
            self.x_prev.masked_fill_(mask.bool(), 0.0)


class VecAgent:
    """``Agent`` (DDPG/DDPG_agent.py:9-52) for ``num_envs`` environments; same constructor arguments."""

    def __init__(self, alpha, beta, input_dims, tau, n_actions, gamma=0.99, max_size=1000000, fc1_dims=400, fc2_dims=300,
                 batch_size=64, num_envs=1, device=None, seed=27, global_env_offset=0, precision="auto", actor_seed=None,
                 allow_out_of_bar=False):
        _lib.require_cuda()
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.gamma, self.tau, self.batch_size, self.alpha, self.beta = gamma, tau, batch_size, alpha, beta
        in_dim = int(input_dims[0]) if isinstance(input_dims, (tuple, list)) else int(input_dims)
        if n_actions != 1:
            raise ValueError("the CUDA actor is specialised to n_actions == 1 (simv2 action space)")
        self.num_envs = int(num_envs)
        self.allow_out_of_bar = bool(allow_out_of_bar)
        self.precision, _ = _lib.resolve_precision(precision, self.allow_out_of_bar)
        self.memory = DeviceReplayBuffer(max_size, (in_dim,), n_actions, device=self.device)
        self.noise = OUNoiseState(self.num_envs, self.device, seed, global_env_offset)
        self.actor = CudaActor(in_dim, fc1_dims, fc2_dims, device=self.device)
        self.actor.load_state_dict(init_actor_state_dict(in_dim, fc1_dims, fc2_dims, n_actions, seed=actor_seed))
        self._action = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        self._learner = None

    def choose_action(self, observation, evaluate=False):
        """DDPG_agent.py:36-49: mu(obs) + OU noise (unless ``evaluate``); returns the UNCLIPPED action [N,1].  One launch
        (``tt_actor_choose_action``: the noise is added in the actor kernel's output stage).  ``observation`` must hold one row
        per environment: the OU state is per env."""
        with torch.cuda.device(self.device):
            obs = torch.as_tensor(observation, device=self.device)
            single = obs.dim() == 1
            if single:
                obs = obs.reshape(1, -1)
            if obs.shape[0] != self.num_envs:
                raise ValueError(f"choose_action: got {obs.shape[0]} observation rows for {self.num_envs} environments")
            if obs.dtype != torch.float32 or obs.stride(1) != 1:
                obs = obs.to(torch.float32).contiguous()
            n = self.noise
            _, prec = _lib.resolve_precision(self.precision, self.allow_out_of_bar)
            check(self.L.tt_actor_choose_action(self.actor._h, obs.data_ptr(), obs.stride(0), self.num_envs, n.x_prev.data_ptr(), n.seed,
                                                n.global_env_offset, n.iter_ptr, int(evaluate), self._action.data_ptr(), None, prec,
                                                None, stream_ptr()))
            if not evaluate and n._own_iter:
                n._iter += 1
            mu = self._action
        return mu.reshape(-1) if single else mu.reshape(-1, 1)

    @staticmethod
    def scale_action(action, high=0.78539819):
        """trainv2.py:516: clip(action, -1, 1) * env.action_space.high"""
        return torch.clamp(action, -1.0, 1.0) * high

    def remember(self, state, action, reward, state_, done):
        """DDPG_agent.py:51-52 for a batch of transitions."""
        self.memory.store_transition(state, action, reward, state_, done)

    def load_actor_state_dict(self, sd):
        self.actor.load_state_dict(sd)

    # ---- checkpoint interop with the reference (DDPG_agent.py:54-70, networks.py:69-95,149-169) ----
    chkpt_dir = "tmp/ddpg"

    def _state_dicts(self):
        if self._learner is not None:                            # the learner owns the four networks once it exists
            return {n: self._learner.state_dict(n) for n in ("actor", "target_actor", "critic", "target_critic")}
        sd = self.actor.state_dict()
        return {"actor": sd, "target_actor": sd}                 # update_network_parameters(tau=1), DDPG_agent.py:34

    def save_models(self):
        """``Agent.save_models``: <chkpt_dir>/{actor,target_actor,critic,target_critic}_ddpg in the reference's format
        (the critics exist once ``learn()`` has run)."""
        from . import checkpoint
        return checkpoint.save_models(self._state_dicts(), self.chkpt_dir)

    def save_models_progress(self, success):
        from . import checkpoint
        return checkpoint.save_models(self._state_dicts(), self.chkpt_dir, progress=success)

    def load_models(self):
        """``Agent.load_models``: loads whatever of the four reference checkpoints exists; the actor goes to the CUDA
        kernels' weight layouts (fp32 + tensor-core images), all four to the learner."""
        from . import checkpoint
        sds = checkpoint.load_models(self.chkpt_dir)
        if "actor" not in sds:
            raise FileNotFoundError(checkpoint.checkpoint_path(self.chkpt_dir, "actor"))
        checkpoint.check_actor_state_dict(sds["actor"], *self.actor.dims)
        self.actor.load_state_dict(sds["actor"])
        if len(sds) > 1 or self._learner is not None:
            ln = self.learner
            for name in ("actor", "target_actor", "critic", "target_critic"):
                if name in sds:
                    ln.load_state_dict(name, sds[name])
            ln.push_actor()
        return sorted(sds)

    @property
    def learner(self):
        """The ``CudaLearner`` (created on first use: critic / targets initialised like the reference ``Agent.__init__``)."""
        if self._learner is None:
            from .learner import CudaLearner
            self._learner = CudaLearner(self)
        return self._learner

    def learn(self):
        """``Agent.learn`` (DDPG_agent.py:72-106): one DDPG update on a batch sampled from the device ring; the new policy is
        re-packed into the rollout actor.  Hand-written kernels (``tt_learn_step``), no torch autograd."""
        return self.learner.learn()


class Agent(VecAgent):
    """N=1 numpy-facing agent with the reference's exact call contract (DDPG_agent.py:36-52):
    ``choose_action(obs(23,)) -> np.float32 (1,)``; ``remember`` takes numpy/python scalars."""

    def __init__(self, alpha, beta, input_dims, tau, n_actions, gamma=0.99, max_size=1000000, fc1_dims=400, fc2_dims=300,
                 batch_size=64, **kw):
        super().__init__(alpha, beta, input_dims, tau, n_actions, gamma, max_size, fc1_dims, fc2_dims, batch_size, num_envs=1, **kw)

    # host <-> device staging of the N = 1 contract: pinned buffers, ONE copy per direction and call (pageable copies of a few
    # bytes each were most of the 123 us of `choose_action` and of the 5-copy `remember`)
    def _staging(self):
        st = getattr(self, "_stage", None)
        if st is None:
            d = self.actor.dims[0]
            with torch.cuda.device(self.device):
                st = self._stage = dict(
                    obs_h=torch.zeros(d, dtype=torch.float32).pin_memory(), obs_d=torch.zeros(1, d, dtype=torch.float32, device=self.device),
                    act_h=torch.zeros(1, dtype=torch.float32).pin_memory(),
                    # one transition: s [d] | a | r | s' [d] as float32, then done as one byte
                    tr_h=torch.zeros(4 * (2 * d + 2) + 4, dtype=torch.uint8).pin_memory(),
                    tr_d=torch.zeros(4 * (2 * d + 2) + 4, dtype=torch.uint8, device=self.device))
                st["tr_hf"] = st["tr_h"].numpy()[:4 * (2 * d + 2)].view(np.float32)
                f = st["tr_d"][:4 * (2 * d + 2)].view(torch.float32)
                st["tr_views"] = (f[:d].view(1, d), f[d:d + 1], f[d + 1:d + 2], f[d + 2:2 * d + 2].view(1, d), st["tr_d"][4 * (2 * d + 2):4 * (2 * d + 2) + 1])
        return st

    def choose_action(self, observation, evaluate=False):
        st = self._staging()
        st["obs_h"].numpy()[:] = np.asarray(observation, np.float32).reshape(-1)
        with torch.cuda.device(self.device):
            st["obs_d"].copy_(st["obs_h"].view(1, -1), non_blocking=True)
            a = super().choose_action(st["obs_d"], evaluate)
            st["act_h"].copy_(a.reshape(-1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return st["act_h"].numpy().copy()

    def remember(self, state, action, reward, state_, done):
        st = self._staging()
        d, h = self.actor.dims[0], st["tr_hf"]
        h[:d] = np.asarray(state, np.float32).reshape(-1)
        h[d] = np.float32(np.asarray(action, np.float32).reshape(-1)[0])
        h[d + 1] = np.float32(reward)
        h[d + 2:2 * d + 2] = np.asarray(state_, np.float32).reshape(-1)
        st["tr_h"].numpy()[4 * (2 * d + 2)] = 1 if done else 0
        with torch.cuda.device(self.device):
            st["tr_d"].copy_(st["tr_h"], non_blocking=True)
            s0, a0, r0, s1, d0 = st["tr_views"]
            super().remember(s0, a0, r0, s1, d0)
            torch.cuda.current_stream().synchronize()          # the staging buffers are reused by the next call
