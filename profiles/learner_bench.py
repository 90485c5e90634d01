"""Timing of the native learner step (tt_learn_step): eager and as one CUDA graph, with and without the policy re-pack.
    python profiles/learner_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpg_trucktrailer_b200 as tt

cap, B = 1 << 16, 64
ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=cap, batch_size=B, num_envs=cap, precision="f16", actor_seed=0)
u = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
ag.remember(u(cap, 23), u(cap), u(cap) * 20, u(cap, 23), (u(cap) > 0.9).to(torch.uint8))
ln = ag.learner
L = tt.load()


def timed(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.tt_launch_count()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (L.tt_launch_count() - l0) / n


for name, kw in (("with re-pack", dict(repack_into="agent")), ("without re-pack", dict(repack_into=None))):
    us, nl = timed(lambda: ln.learn(**kw))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ln.learn(**kw)
    usg, _ = timed(g.replay)
    print(f"learner step {name}: eager {us:.1f} us ({nl:.0f} launches), one CUDA graph {usg:.1f} us")
