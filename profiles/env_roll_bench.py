"""The rollout's env kernel (tt_env_step_reset: step + ring store + deferred in-kernel reset + tick) against its parts, each
timed alone with CUDA events at N = 2^22 after a pre-roll that brings episodes to their stationary mix:
    a) tt_env_step                (no ring)            b) tt_env_step_store (ring: s', r, done)
    c) tt_env_step_reset, no ring                     d) tt_env_step_reset with ring (the rollout's launch)
    e) tt_env_reset(mask = done) + tt_env_tick (what c/d replace)
    python profiles/env_roll_bench.py [N]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
cap = 1 << 24
L = tt.load(); s = _lib.stream_ptr()
env = tt.VecTruckTrailerEnv(N, seed=27, auto_tick=False)
obs, _ = env.reset()
mem = tt.DeviceReplayBuffer(cap)
x = torch.zeros(N, device="cuda")
o2 = torch.zeros(N, 23, device="cuda"); rew = torch.zeros(N, device="cuda"); done = torch.zeros(N, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = [torch.empty(N, device="cuda").uniform_(-0.6, 0.6, generator=g) for _ in range(8)]
cntr = [0]
def ring():
    r = _lib.ReplayRing(mem.state_memory.data_ptr(), mem.action_memory.data_ptr(), mem.reward_memory.data_ptr(),
                        mem.new_state_memory.data_ptr(), mem.terminal_memory.data_ptr(), cap, cntr[0])
    cntr[0] += N
    return r
def roll(i, with_ring=True):
    r = ring()
    _lib.check(L.tt_env_step_reset(env._h, acts[i % 8].data_ptr(), o2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), x.data_ptr(), C.byref(r) if with_ring else None, s))
for i in range(200):
    roll(i)
torch.cuda.synchronize()
print("finished per step: %.2f %%" % (100.0 * float(done.float().mean())))

def timed(fn, reps=20):
    tot = 0.0
    for i in range(reps):
        roll(i)                                     # keep the population stationary between samples
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3

def a(i): _lib.check(L.tt_env_step(env._h, acts[i % 8].data_ptr(), o2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), None, s))
def b(i):
    r = ring(); _lib.check(L.tt_env_step_store(env._h, acts[i % 8].data_ptr(), o2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), C.byref(r), s))
def c(i): roll(i, False)
def d(i): roll(i, True)
def e(i):
    _lib.check(L.tt_env_reset(env._h, done.data_ptr(), o2.data_ptr(), 23, s)); _lib.check(L.tt_env_tick(env._h, 1, s))
for name, fn, byt in (("a) step", a, 229), ("b) step + ring", b, 326), ("c) step + reset + tick", c, 229), ("d) step + ring + reset + tick", d, 326), ("e) reset kernel + tick kernel", e, 0)):
    us = timed(fn)
    # a) and b) leave finished envs frozen: reset them so that the next sample sees the same population
    print(f"{name:34s} {us:7.1f} us" + (f"   {N * byt / us / 1e3:7.0f} GB/s = {N * byt / us / 1e3 / 6553.3 * 100:.1f} % of 6553 GB/s" if byt else ""))
