"""Learner step of the reference (``Agent.learn``, DDPG/DDPG_agent.py:72-131; critic DDPG/networks.py:9-68)
as plain PyTorch on the device-resident replay ring.  NOT part of the B200 hot path (SURVEY.md section 8f,
row f1 "next"): it exists so that a driver written against the reference API runs end to end, and it
hands updated actor weights back to the CUDA actor.

``TorchLearner(agent, graph=True)`` captures the whole step -- index sampling, the ring gather kernel (tt_replay_gather),
both forward / backward passes, Adam, the soft target update and the re-pack of the new actor weights into the CUDA
actor's fp32 / tensor-core layouts (tt_actor_load) -- into ONE CUDA graph (fused Adam, multi-tensor soft update): the step is
~200 tiny launches, i.e. pure launch latency in eager mode (3.0 ms eager, 0.51 ms as a graph on a B200)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .agent import ACTOR_KEYS


class _Actor(nn.Module):
    def __init__(self, i, h1, h2):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(i, h1), nn.Linear(h1, h2)
        self.bn1, self.bn2 = nn.LayerNorm(h1), nn.LayerNorm(h2)
        self.mu = nn.Linear(h2, 1)

    def forward(self, s):
        x = F.relu(self.bn1(self.fc1(s)))
        x = F.relu(self.bn2(self.fc2(x)))
        return torch.tanh(self.mu(x))


class _Critic(nn.Module):
    def __init__(self, i, h1, h2):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(i, h1), nn.Linear(h1, h2)
        self.bn1, self.bn2 = nn.LayerNorm(h1), nn.LayerNorm(h2)
        self.action_value, self.q = nn.Linear(1, h2), nn.Linear(h2, 1)
        for lin, f in ((self.fc1, 1 / np.sqrt(h1)), (self.fc2, 1 / np.sqrt(h2)), (self.q, 0.003), (self.action_value, 1 / np.sqrt(h2))):
            lin.weight.data.uniform_(-f, f); lin.bias.data.uniform_(-f, f)

    def forward(self, s, a):
        x = F.relu(self.bn1(self.fc1(s)))
        x = self.bn2(self.fc2(x))
        return self.q(F.relu(x + self.action_value(a)))


class TorchLearner:
    def __init__(self, agent, graph=False):
        self.agent = agent
        self.use_graph = bool(graph)
        self._graph = None
        i, h1, h2 = agent.actor.dims
        dev = agent.device
        self.actor, self.target_actor = _Actor(i, h1, h2).to(dev), _Actor(i, h1, h2).to(dev)
        self.critic, self.target_critic = _Critic(i, h1, h2).to(dev), _Critic(i, h1, h2).to(dev)
        self.actor.load_state_dict(agent.actor.state_dict())
        self.target_actor.load_state_dict(self.actor.state_dict())
        self.target_critic.load_state_dict(self.critic.state_dict())
        cap = dict(capturable=True, fused=True) if self.use_graph else {}      # one kernel per optimizer step inside the graph
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=agent.alpha, **cap)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=agent.beta, weight_decay=0.01, **cap)

    @torch.no_grad()
    def _soft_update(self, net, target, tau):
        ps, tps = list(net.parameters()), list(target.parameters())
        if self.use_graph:                                       # two multi-tensor kernels instead of two per parameter
            torch._foreach_mul_(tps, 1 - tau)
            torch._foreach_add_(tps, ps, alpha=tau)
            return
        for p, tp in zip(ps, tps):
            tp.mul_(1 - tau).add_(p, alpha=tau)

    def learn(self):
        ag = self.agent
        if ag.memory.mem_cntr < ag.batch_size:
            return
        if self.use_graph:
            return self._learn_graph()
        s, a, r, s2, d = ag.memory.sample_buffer(ag.batch_size)
        self._update(s, a, r, s2, d)
        ag.actor.load_state_dict({k: v.detach() for k, v in self.actor.state_dict().items() if k in ACTOR_KEYS})

    def _update(self, s, a, r, s2, d):
        """DDPG_agent.py:84-106 on one batch."""
        ag = self.agent
        with torch.no_grad():
            q2 = self.target_critic(s2, self.target_actor(s2))
            q2 = q2.masked_fill(d.view(-1, 1), 0.0)                         # critic_value_[done] = 0.0
            target = (r + ag.gamma * q2.view(-1)).view(ag.batch_size, 1)
        self.critic_opt.zero_grad()
        F.mse_loss(target, self.critic(s, a)).backward()
        self.critic_opt.step()
        self.actor_opt.zero_grad()
        (-self.critic(s, self.actor(s))).mean().backward()
        self.actor_opt.step()
        self._soft_update(self.actor, self.target_actor, ag.tau)
        self._soft_update(self.critic, self.target_critic, ag.tau)

    # ------------------------------------------------------------------ the whole step as one CUDA graph
    def _graph_body(self):
        from . import _lib
        ag, m, B = self.agent, self.agent.memory, self.agent.batch_size
        rows = (torch.rand(B, device=ag.device) * self._max_mem).long().clamp_(max=m.mem_size - 1)     # replay_buffer.py:26
        s, a, r, s2, d = self._batch
        _lib.check(m.L.tt_replay_gather(m.state_memory.data_ptr(), m.action_memory.data_ptr(), m.reward_memory.data_ptr(),
                                        m.new_state_memory.data_ptr(), m.terminal_memory.data_ptr(), rows.data_ptr(), B, s.data_ptr(),
                                        a.data_ptr(), r.data_ptr(), s2.data_ptr(), d.data_ptr(), _lib.stream_ptr()))
        self._update(s, a, r, s2, d.bool())
        p = dict(self.actor.named_parameters())
        _lib.check(ag.actor.L.tt_actor_load(ag.actor._h, *[p[k].data_ptr() for k in ACTOR_KEYS], _lib.stream_ptr()))

    def _learn_graph(self):
        ag, m = self.agent, self.agent.memory
        with torch.cuda.device(ag.device):
            if self._graph is None:
                B, dev = ag.batch_size, ag.device
                self._max_mem = torch.zeros((), dtype=torch.float32, device=dev)
                self._batch = (torch.empty(B, 23, device=dev), torch.empty(B, 1, device=dev), torch.empty(B, device=dev),
                               torch.empty(B, 23, device=dev), torch.empty(B, dtype=torch.uint8, device=dev))
                self._max_mem.fill_(float(min(m.mem_cntr, m.mem_size)))
                # the rollout actor reads the learner's parameter tensors from now on (static addresses)
                ag.actor._sd = {k: v for k, v in self.actor.named_parameters() if k in ACTOR_KEYS}
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):                      # warm-up outside the capture (allocator, Adam state, cuBLAS)
                    for _ in range(3):
                        self._graph_body()
                torch.cuda.current_stream().wait_stream(side)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._graph_body()
                self.graph_warmup_steps = 3
                return
            self._max_mem.fill_(float(min(m.mem_cntr, m.mem_size)))
            self._graph.replay()
