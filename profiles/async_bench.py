"""configs[3] (rollout + one DDPG update per iteration) under different schedules, N = 2^22 envs on one GPU:
   rollout alone with r SMs reserved | update serial on the rollout stream | rollout.AsyncTrainer with r SMs reserved.
   python profiles/async_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
L = tt.load()
env = tt.VecTruckTrailerEnv(N, seed=27)
ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=1 << 24, num_envs=N, seed=27, precision="f16", actor_seed=0)
eng = tt.RolloutEngine(env, ag, store=True)
eng.reset()
for _ in range(256):
    eng.step()


def timed(fn, iters=40, warm=60):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for r in (0, 2, 4, 8):
    _lib.check(L.tt_reserve_sms(r))
    print(f"rollout alone, {r} SMs reserved: {timed(eng.step):.4f} ms / iteration")
_lib.check(L.tt_reserve_sms(0))
ln = ag.learner
def serial():
    eng.step(); ln.learn()
print(f"update serial on the rollout stream:  {timed(serial):.4f} ms / iteration")
a0 = ag.actor
for r in (0, 2, 4, 8):
    ag.actor = a0
    tr = tt.AsyncTrainer(eng, reserve_sms=r)
    print(f"AsyncTrainer, {r} SMs reserved:        {timed(tr.step):.4f} ms / iteration  ({tr.updates} updates)")
    tr.close()
