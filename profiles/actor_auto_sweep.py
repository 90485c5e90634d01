"""TT_PREC_AUTO crossover (north_star (c)): latency of one actor forward, warp-level fp32 FMA kernel vs tcgen05 kernel (f16),
by batch size -- back-to-back launches (launch overhead included, as a caller sees it) and inside a CUDA graph.
    python profiles/actor_auto_sweep.py > profiles/r02_actor_auto_sweep.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpg_trucktrailer_b200 as tt

actor = tt.agent.CudaActor(); actor.load_state_dict(tt.init_actor_state_dict(seed=0))


def timed(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("| rows | fp32 eager us | f16 eager us | fp32 in graph us | f16 in graph us | auto picks | max abs diff f16 vs fp32 |")
print("|---|---|---|---|---|---|---|")
for n in (1, 8, 9, 16, 17, 32, 33, 64, 128, 192, 256, 512, 1024, 4096, 16384, 65536):
    obs = torch.empty(n, 23, device="cuda").uniform_(-1, 1)
    out = torch.empty(n, device="cuda")
    row = []
    for prec in ("fp32", "f16"):
        row.append(timed(lambda: actor.forward(obs, out=out, precision=prec)))
    for prec in ("fp32", "f16"):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): actor.forward(obs, out=out, precision=prec)
        row.append(timed(g.replay, 30) / 20)
    d = (actor.forward(obs, precision="f16").clone() - actor.forward(obs, precision="fp32")).abs().max().item()
    print(f"| {n} | {row[0]:.1f} | {row[1]:.1f} | {row[2]:.1f} | {row[3]:.1f} | {actor.auto_precision(n)} | {d:.1e} |")
