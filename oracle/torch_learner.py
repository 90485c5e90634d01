"""TEST INFRASTRUCTURE ONLY -- the learner step of the reference (``Agent.learn``, DDPG/DDPG_agent.py:72-131; critic
DDPG/networks.py:9-68) restated in plain PyTorch (autograd + torch.optim.Adam).  It is pinned against the untouched
reference by tests/test_next_rows_cpu.py::test_learner_step_matches_reference_learn (tests/golden/ref_learn.npz) and serves
as the checker of the hand-written CUDA learner (ddpg-trucktrailer_b200/csrc/tt_learn.cu, tests/test_gpu_learner.py):
per-tensor gradients and parameters after a step.  The product package never imports it."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ACTOR_KEYS = ("fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight", "bn2.bias",
              "mu.weight", "mu.bias")


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
    def __init__(self, agent):
        self.agent = agent
        i, h1, h2 = agent.actor.dims
        dev = agent.device
        self.actor, self.target_actor = _Actor(i, h1, h2).to(dev), _Actor(i, h1, h2).to(dev)
        self.critic, self.target_critic = _Critic(i, h1, h2).to(dev), _Critic(i, h1, h2).to(dev)
        self.actor.load_state_dict(agent.actor.state_dict())
        self.target_actor.load_state_dict(self.actor.state_dict())
        self.target_critic.load_state_dict(self.critic.state_dict())
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=agent.alpha)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=agent.beta, weight_decay=0.01)

    @torch.no_grad()
    def _soft_update(self, net, target, tau):
        ps, tps = list(net.parameters()), list(target.parameters())
        for p, tp in zip(ps, tps):
            tp.mul_(1 - tau).add_(p, alpha=tau)

    def learn(self):
        ag = self.agent
        if ag.memory.mem_cntr < ag.batch_size:
            return
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
