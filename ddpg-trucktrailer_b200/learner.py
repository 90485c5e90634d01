"""Learner step of the reference (``Agent.learn``, DDPG/DDPG_agent.py:72-131; networks DDPG/networks.py:9-68, :98-147;
``ReplayBuffer.sample_buffer`` DDPG/replay_buffer.py:23-34) on the device-resident replay ring: row f1 of SURVEY.md
section 8.  ``CudaLearner`` binds ``tt_learn_step`` (csrc/tt_learn.cu): 14 hand-written kernels -- sampling + gather, the
forward passes of the four networks as grouped launches, both backward passes, Adam (critic weight decay 0.01), the soft
target updates -- followed by the re-pack of the new policy into the rollout actor's operand images.  No torch.nn, no
autograd, no cuBLAS; every launch goes to the caller's stream and the sequence is graph-capturable.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, stream_ptr

ACTOR_KEYS = ("fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight", "bn2.bias",
              "mu.weight", "mu.bias")
# the reference CriticNetwork's state_dict order (module registration order, networks.py:20-33)
CRITIC_KEYS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias",
               "action_value.weight", "action_value.bias", "q.weight", "q.bias")
NETS = ("actor", "target_actor", "critic", "target_critic")


def _flat_layout(kind, i, h1, h2):
    """name -> (offset, shape) inside the flat float32 parameter vector of one network (include/tt_b200.h)."""
    trunk = [("fc1.weight", (h1, i)), ("fc1.bias", (h1,)), ("bn1.weight", (h1,)), ("bn1.bias", (h1,)),
             ("fc2.weight", (h2, h1)), ("fc2.bias", (h2,)), ("bn2.weight", (h2,)), ("bn2.bias", (h2,))]
    tail = [("mu.weight", (1, h2)), ("mu.bias", (1,))] if kind == "actor" else \
           [("action_value.weight", (h2, 1)), ("action_value.bias", (h2,)), ("q.weight", (1, h2)), ("q.bias", (1,))]
    out, off = {}, 0
    for name, shape in trunk + tail:
        out[name] = (off, shape)
        off += int(np.prod(shape))
    return out, off


def init_critic_state_dict(input_dims=23, fc1_dims=400, fc2_dims=300, n_actions=1, seed=None):
    """Initial critic parameters with the reference's distributions and draw order (networks.py:20-45): fc1 / fc2
    U(+-1/sqrt(out_features)), q U(+-0.003), action_value U(+-1/sqrt(fc2_dims)), LayerNorm (1, 0)."""
    import torch.nn as nn
    if seed is not None:
        torch.manual_seed(seed)
    fc1, fc2 = nn.Linear(input_dims, fc1_dims), nn.Linear(fc1_dims, fc2_dims)
    bn1, bn2 = nn.LayerNorm(fc1_dims), nn.LayerNorm(fc2_dims)
    av, q = nn.Linear(n_actions, fc2_dims), nn.Linear(fc2_dims, 1)
    for lin, f in ((fc1, 1.0 / np.sqrt(fc1_dims)), (fc2, 1.0 / np.sqrt(fc2_dims)), (q, 0.003), (av, 1.0 / np.sqrt(fc2_dims))):
        lin.weight.data.uniform_(-f, f); lin.bias.data.uniform_(-f, f)
    mods = {"fc1": fc1, "fc2": fc2, "bn1": bn1, "bn2": bn2, "action_value": av, "q": q}
    return {k: getattr(mods[k.split(".")[0]], k.split(".")[1]).data.clone() for k in CRITIC_KEYS}


class CudaLearner:
    """The four networks of the reference ``Agent`` + both Adam states + the batch workspace in one device workspace;
    ``learn()`` = one ``Agent.learn()``.  Parameters are exchanged as reference-layout state_dicts."""

    def __init__(self, agent, seed=None, critic_seed=None, critic_weight_decay=0.01):
        _lib.require_cuda()
        self.agent = agent
        self.L = _lib.load()
        self.device = agent.device
        i, h1, h2 = agent.actor.dims
        self.dims, self.batch = (i, h1, h2), int(agent.batch_size)
        with torch.cuda.device(self.device):
            nbytes = self.L.tt_learner_workspace_bytes(i, h1, h2, self.batch)
            if nbytes == 0:
                raise ValueError("learner: unsupported sizes")
            self._ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 255) // 256 * 256
            h = C.c_void_p()
            check(self.L.tt_learner_create(C.byref(h), i, h1, h2, self.batch, float(agent.alpha), float(agent.beta), float(agent.gamma),
                                           float(agent.tau), float(critic_weight_decay), int(agent.noise.seed if seed is None else seed),
                                           base, nbytes))
            self._h = h
            # float32 views of the flat parameter / gradient vectors inside the workspace
            self._flat, self._params, self._gflat = {}, {}, {}
            for n, name in enumerate(NETS):
                kind = "actor" if "actor" in name else "critic"
                layout, count = _flat_layout(kind, i, h1, h2)
                assert count == self.L.tt_learner_param_count(h, n)
                off = self.L.tt_learner_params(h, n) - self._ws.data_ptr()
                flat = self._ws[off:off + 4 * count].view(torch.float32)
                self._flat[name] = flat
                self._params[name] = {k: flat[o:o + int(np.prod(shp))].view(shp) for k, (o, shp) in layout.items()}
            for which, name in enumerate(("actor", "critic")):
                off = self.L.tt_learner_grads(h, which) - self._ws.data_ptr()
                self._gflat[name] = self._ws[off:off + 4 * self._flat[name].numel()].view(torch.float32)
            # Agent.__init__ (DDPG_agent.py:22-34): targets start as copies (update_network_parameters(tau=1))
            self.load_state_dict("actor", agent.actor.state_dict())
            self.load_state_dict("target_actor", agent.actor.state_dict())
            csd = init_critic_state_dict(i, h1, h2, 1, seed=critic_seed)
            self.load_state_dict("critic", csd)
            self.load_state_dict("target_critic", csd)
        self.steps = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.L.tt_learner_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- parameter exchange (reference state_dict layout) ----
    def state_dict(self, net):
        keys = ACTOR_KEYS if "actor" in net else CRITIC_KEYS
        return {k: self._params[net][k].detach().clone() for k in keys}

    def load_state_dict(self, net, sd):
        for k, view in self._params[net].items():
            t = torch.as_tensor(sd[k]).detach().to(device=self.device, dtype=torch.float32)
            if tuple(t.shape) != tuple(view.shape):
                raise ValueError(f"{net}.{k}: expected shape {tuple(view.shape)}, got {tuple(t.shape)}")
            view.copy_(t)

    def grads(self, net):
        """Gradients of the last step (``actor`` or ``critic``) as a name -> tensor dict (diagnostics / tests)."""
        kind = "actor" if net == "actor" else "critic"
        layout, _ = _flat_layout(kind, *self.dims)
        g = self._gflat[kind]
        return {k: g[o:o + int(np.prod(shp))].view(shp).clone() for k, (o, shp) in layout.items()}

    def reset_optimizer(self):
        with torch.cuda.device(self.device):
            check(self.L.tt_learner_reset_optimizer(self._h, stream_ptr()))

    def push_actor(self, actor=None):
        """Re-pack the learner's current actor parameters into a rollout actor (default: the agent's)."""
        actor = actor or self.agent.actor
        actor.load_state_dict(self._params["actor"])
        actor._sd = {k: self._params["actor"][k] for k in ACTOR_KEYS}       # the rollout actor's state_dict follows the learner

    # ---- Agent.learn ----
    def learn(self, rows=None, repack_into="agent", window=None):
        """One DDPG update on a batch sampled from the agent's ring (``rows``: explicit ring rows instead, int64 [batch]).
        ``repack_into``: rollout actor that receives the new policy ("agent" = the agent's, None = nobody).
        ``window = (begin, count)``: sample only from `count` ring rows starting at `begin` (wrapping) -- for a learner
        that runs concurrently with a rollout iteration (``rollout.AsyncTrainer``)."""
        ag, m = self.agent, self.agent.memory
        if (m.mem_cntr if window is None else window[1]) < self.batch:      # DDPG_agent.py:73-74
            return
        with torch.cuda.device(self.device):
            ring = _lib.ReplayRing(m.state_memory.data_ptr(), m.action_memory.data_ptr(), m.reward_memory.data_ptr(),
                                   m.new_state_memory.data_ptr(), m.terminal_memory.data_ptr(), m.mem_size, m.mem_cntr)
            if rows is not None:
                rows = torch.as_tensor(rows, dtype=torch.int64, device=self.device).contiguous()
                if rows.numel() != self.batch:
                    raise ValueError(f"rows: expected {self.batch} ring rows")
            target = ag.actor if repack_into == "agent" else repack_into
            if window is None:
                check(self.L.tt_learn_step(self._h, C.byref(ring), None if rows is None else rows.data_ptr(),
                                           target._h if target is not None else None, stream_ptr()))
            else:
                check(self.L.tt_learn_step_window(self._h, C.byref(ring), None if rows is None else rows.data_ptr(),
                                                  target._h if target is not None else None, int(window[0]), int(window[1]), stream_ptr()))
            if target is not None:
                target._sd = {k: self._params["actor"][k] for k in ACTOR_KEYS}
        self.steps += 1
