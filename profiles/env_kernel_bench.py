"""Micro-benchmark of the env step kernel alone (CUDA events, N = 2^22): single-step launches with the reset
kernel keeping the population alive, and K-step launches (state in registers)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddpg_trucktrailer_b200 as tt

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
env = tt.VecTruckTrailerEnv(N, seed=27)
env.reset()
act = torch.empty(N, device="cuda").uniform_(-0.6, 0.6)
def one():
    _, r, d, _ = env.step(act)
    return d
for _ in range(5):
    d = one(); env.reset(options={"mask": d}); env.tick()
tot = 0.0
for _ in range(30):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d = one(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
    env.reset(options={"mask": d}); env.tick()
ms = tot / 30
print(f"MINBLOCKS={os.environ.get('TT_ENV_MINBLOCKS','default')} N={N} env_step {ms*1e3:.1f} us  {N/ms/1e6:.1f} Gsteps/s  algorithmic {N*229/ms/1e6:.0f} GB/s = {N*229/ms/1e6/6553.3*100:.1f}% of 6553 GB/s")
K = 16
acts = torch.empty(K, N, device="cuda").uniform_(-0.6, 0.6)
env.step_k(acts, auto_reset=True, want_reward=False, want_done=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    env.step_k(acts, auto_reset=True, want_reward=False, want_done=False)
e1.record(); torch.cuda.synchronize()
print(f"   step_k K={K} auto-reset, no per-step outputs: {e0.elapsed_time(e1)/3/K*1e3:.1f} us per step")
