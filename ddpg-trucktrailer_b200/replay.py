"""Device-resident replay ring: ``ReplayBuffer`` of DDPG/replay_buffer.py:4-34 with batched stores.

Same attribute names as the reference (``mem_size, mem_cntr, state_memory, new_state_memory, action_memory,
reward_memory, terminal_memory``); the arrays are CUDA float32 (the reference keeps float64 numpy arrays
of float32 payloads).  ``store_transition`` takes a whole batch of transitions and is equivalent to that
many sequential reference calls in env order (row ``(mem_cntr + i) % mem_size``, last writer wins).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import TT_OBS_DIM, check, stream_ptr


class DeviceReplayBuffer:
    def __init__(self, max_size, input_shape=(TT_OBS_DIM,), n_actions=1, device=None, seed=0):
        _lib.require_cuda()
        if tuple(input_shape) != (TT_OBS_DIM,) or n_actions != 1:
            raise ValueError("the CUDA replay ring is specialised to 23-dim observations and 1 action")
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.mem_size = int(max_size)
        self.mem_cntr = 0
        dev = self.device
        self.state_memory = torch.zeros(self.mem_size, TT_OBS_DIM, dtype=torch.float32, device=dev)
        self.new_state_memory = torch.zeros(self.mem_size, TT_OBS_DIM, dtype=torch.float32, device=dev)
        self.action_memory = torch.zeros(self.mem_size, 1, dtype=torch.float32, device=dev)
        self.reward_memory = torch.zeros(self.mem_size, dtype=torch.float32, device=dev)
        self.terminal_memory = torch.zeros(self.mem_size, dtype=torch.uint8, device=dev)
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(seed)

    def store_transition(self, state, action, reward, state_, done):
        """replay_buffer.py:13-21 for a batch: state/state_ [n,23] (row stride may exceed 23), action [n] or
        [n,1], reward [n], done [n] (bool or uint8)."""
        with torch.cuda.device(self.device):
            s, s2 = _rows(state), _rows(state_)
            n = s.shape[0]
            a = action.reshape(-1).contiguous()
            r = reward.reshape(-1).contiguous()
            d = done.reshape(-1)
            d = (d.to(torch.uint8) if d.dtype != torch.uint8 else d).contiguous()
            check(self.L.tt_replay_store(self.state_memory.data_ptr(), self.action_memory.data_ptr(),
                                         self.reward_memory.data_ptr(), self.new_state_memory.data_ptr(),
                                         self.terminal_memory.data_ptr(), self.mem_size, self.mem_cntr,
                                         s.data_ptr(), s.stride(0), a.data_ptr(), r.data_ptr(), s2.data_ptr(), s2.stride(0),
                                         d.data_ptr(), n, stream_ptr()))
            self.mem_cntr += n

    def sample_buffer(self, batch_size):
        """replay_buffer.py:23-34: uniform with replacement over the filled part."""
        with torch.cuda.device(self.device):
            max_mem = min(self.mem_cntr, self.mem_size)
            rows = torch.randint(0, max_mem, (batch_size,), device=self.device, generator=self._gen, dtype=torch.int64)
            dev = self.device
            s = torch.empty(batch_size, TT_OBS_DIM, dtype=torch.float32, device=dev)
            s2 = torch.empty_like(s)
            a = torch.empty(batch_size, 1, dtype=torch.float32, device=dev)
            r = torch.empty(batch_size, dtype=torch.float32, device=dev)
            d = torch.empty(batch_size, dtype=torch.uint8, device=dev)
            check(self.L.tt_replay_gather(self.state_memory.data_ptr(), self.action_memory.data_ptr(),
                                          self.reward_memory.data_ptr(), self.new_state_memory.data_ptr(),
                                          self.terminal_memory.data_ptr(), rows.data_ptr(), batch_size, s.data_ptr(),
                                          a.data_ptr(), r.data_ptr(), s2.data_ptr(), d.data_ptr(), stream_ptr()))
        return s, a, r, s2, d.bool()


def _rows(t):
    """[n,23] float32 CUDA view with unit inner stride (row stride may be larger)."""
    if t.dim() == 1:
        t = t.reshape(1, -1)
    if t.dtype != torch.float32 or t.stride(1) != 1:
        t = t.to(torch.float32).contiguous()
    return t


ReplayBuffer = DeviceReplayBuffer
