"""Replay ring scatter / gather kernels vs the reference's ReplayBuffer semantics (bit-exact)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(rng, n, ld=23):
    s = torch.zeros(n, ld); s2 = torch.zeros(n, ld)
    s[:, :23] = torch.from_numpy(rng.uniform(-1, 1, (n, 23)).astype(np.float32))
    s2[:, :23] = torch.from_numpy(rng.uniform(-1, 1, (n, 23)).astype(np.float32))
    a = torch.from_numpy(rng.uniform(-1.3, 1.3, n).astype(np.float32))
    r = torch.from_numpy(rng.normal(0, 100, n).astype(np.float32))
    d = torch.from_numpy((rng.random(n) < 0.2).astype(np.uint8))
    return s, a, r, s2, d


def test_reference_wraparound_fixture(golden_dir):
    """replay_buffer.py:13-21: capacity 10, 27 stores (wraps twice), recorded from the reference class."""
    import ddpg_trucktrailer_b200 as tt
    g = np.load(os.path.join(golden_dir, "ref_misc.npz"))
    rb = tt.DeviceReplayBuffer(10)
    cu = lambda x, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(x)).to(dt).cuda()
    s, s2, a, r, d = cu(g["rb_s"]), cu(g["rb_s2"]), cu(g["rb_a"][:, 0]), cu(g["rb_r"]), cu(g["rb_d"], torch.uint8)
    for lo, hi in ((0, 4), (4, 5), (5, 20), (20, 27)):          # ragged batch sizes, one of them wraps
        rb.store_transition(s[lo:hi], a[lo:hi], r[lo:hi], s2[lo:hi], d[lo:hi])
    assert rb.mem_cntr == int(g["rb_cntr"]) == 27
    assert np.array_equal(rb.state_memory.cpu().numpy().astype(np.float64), g["rb_state"])
    assert np.array_equal(rb.new_state_memory.cpu().numpy().astype(np.float64), g["rb_new_state"])
    assert np.array_equal(rb.action_memory.cpu().numpy().astype(np.float64), g["rb_action"])
    assert np.array_equal(rb.reward_memory.cpu().numpy(), g["rb_reward"].astype(np.float32))
    assert np.array_equal(rb.terminal_memory.cpu().numpy().astype(bool), g["rb_terminal"])


@pytest.mark.parametrize("cap,sizes,ld", [(1000, [1, 7, 333, 659, 5, 1000, 2], 23), (4096, [4096, 100, 8000, 3], 24),
                                          (777, [3000, 1], 23), (1 << 16, [1 << 16, 12345], 23)])
def test_store_equals_sequential_semantics(cap, sizes, ld):
    """n sequential store_transition calls == one batched store (incl. n > mem_size: last writer wins)."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    rng = np.random.default_rng(cap)
    rb = tt.DeviceReplayBuffer(cap)
    S = np.zeros((cap, 23), np.float32); S2 = np.zeros((cap, 23), np.float32)
    A = np.zeros(cap, np.float32); R = np.zeros(cap, np.float32); D = np.zeros(cap, np.uint8)
    cntr = 0
    for n in sizes:
        s, a, r, s2, d = _batch(rng, n, ld)
        rb.store_transition(s.cuda()[:, :23], a.cuda(), r.cuda(), s2.cuda()[:, :23], d.cuda())
        orc.replay_store(S, A, R, S2, D, cntr, s.numpy(), a.numpy(), r.numpy(), s2.numpy(), d.numpy())
        cntr += n
    assert rb.mem_cntr == cntr
    assert np.array_equal(rb.state_memory.cpu().numpy(), S) and np.array_equal(rb.new_state_memory.cpu().numpy(), S2)
    assert np.array_equal(rb.action_memory.cpu().numpy()[:, 0], A) and np.array_equal(rb.reward_memory.cpu().numpy(), R)
    assert np.array_equal(rb.terminal_memory.cpu().numpy(), D)


def test_sample_buffer_gathers_stored_rows():
    """replay_buffer.py:23-34: every sampled tuple is a stored transition; only the filled part is sampled."""
    import ddpg_trucktrailer_b200 as tt
    rng = np.random.default_rng(0)
    rb = tt.DeviceReplayBuffer(5000)
    s, a, r, s2, d = _batch(rng, 1200)
    rb.store_transition(s.cuda(), a.cuda(), r.cuda(), s2.cuda(), d.cuda())
    bs, ba, br, bs2, bd = rb.sample_buffer(256)
    assert bs.shape == (256, 23) and ba.shape == (256, 1) and bd.dtype == torch.bool
    key = {float(x): i for i, x in enumerate(r.numpy())}
    for j in range(256):
        i = key[float(br[j])]
        assert torch.equal(bs[j].cpu(), s[i]) and torch.equal(bs2[j].cpu(), s2[i]) and float(ba[j]) == float(a[i]) and bool(bd[j]) == bool(d[i])
