"""One-call rollout iteration: actor -> OU noise -> clip*pi/4 -> env step -> replay store -> reset of finished
envs, i.e. the body of the reference training loop (DDPG/trainv2.py:511-531) without ``learn()``, for N
environments, issued as one launch sequence through ``tt_rollout_step``."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import PRECISIONS, check, ptr, stream_ptr


class RolloutEngine:
    def __init__(self, env, agent, store=True, evaluate=False, precision=None):
        self.env, self.agent = env, agent
        self.store, self.evaluate = store, evaluate
        self.precision = precision or agent.precision
        self.L = _lib.load()
        agent.noise.bind_env(env)
        N, dev = env.num_envs, env.device
        self.action = torch.zeros(N, dtype=torch.float32, device=dev)
        self.scaled = torch.zeros(N, dtype=torch.float32, device=dev)
        self.reward = env._reward
        self.done = env._done
        self.iterations = 0

    def reset(self, seed=None):
        obs, _ = self.env.reset(seed=seed)
        self.agent.noise.bind_env(self.env)
        self.agent.noise.reset()
        return obs

    def step(self):
        """One iteration for all envs.  Returns (obs_next, reward, done) views of device buffers; finished envs'
        rows of obs_next already hold the reset observation (their terminal observation went to the ring)."""
        env, ag = self.env, self.agent
        with torch.cuda.device(env.device):
            cur = env._obs[env._cur]
            env._cur ^= 1
            nxt = env._obs[env._cur]
            m = ag.memory
            b = _lib.RolloutBufs(cur.data_ptr(), nxt.data_ptr(), env.ld_obs, ag.noise.x_prev.data_ptr(), self.action.data_ptr(),
                                 self.scaled.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                                 m.state_memory.data_ptr() if self.store else None,
                                 m.action_memory.data_ptr() if self.store else None,
                                 m.reward_memory.data_ptr() if self.store else None,
                                 m.new_state_memory.data_ptr() if self.store else None,
                                 m.terminal_memory.data_ptr() if self.store else None, m.mem_size, m.mem_cntr)
            prec = PRECISIONS[self.precision]
            check(self.L.tt_rollout_step(env._h, ag.actor._h, C.byref(b), prec, int(self.evaluate), stream_ptr()))
            if self.store:
                m.mem_cntr += env.num_envs
            self.iterations += 1
        return env._obs_view(nxt), self.reward, self.done
