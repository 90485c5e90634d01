"""Marginal cost of each of the learner's 14 launches in the warm, pipelined state: a CUDA graph of the first k launches for
k = 1..14 (+ the two re-pack launches), differences of consecutive graph times.  Needs the developer knobs TT_LEARN_STOP /
TT_PACK_STOP, which only a -DTT_LEARN_PROFILE build has:
    TT_NVCC_EXTRA=-DTT_LEARN_PROFILE python ddpg-trucktrailer_b200/build.py --force && python profiles/learner_stage_times.py
(rebuild without the flag afterwards)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpg_trucktrailer_b200 as tt

cap, B = 1 << 16, 64
ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=cap, batch_size=B, num_envs=cap, precision="f16", actor_seed=0)
u = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
ag.remember(u(cap, 23), u(cap), u(cap) * 20, u(cap, 23), (u(cap) > 0.9).to(torch.uint8))
ln = ag.learner
names = ["sampling + gather + fc1 x4", "fc2 x4 (gemm)", "critic head", "critic dW2+da1 (gemm)", "critic LN1 backward", "critic dW1 (gemm)", "Adam critic",
         "fc1 critic", "fc2 critic (gemm)", "actor head", "actor dW2+da1 (gemm)", "actor LN1 backward", "actor dW1 (gemm)", "Adam actor"]


def timed(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


prev = 0.0
for k in range(1, 15):
    os.environ["TT_LEARN_STOP"] = str(k)
    ln.learn(repack_into=None)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ln.learn(repack_into=None)
    t = timed(g.replay)
    print(f"{k:2d} {names[k - 1]:28s} cumulative {t:7.1f} us   this launch {t - prev:6.1f} us")
    prev = t

os.environ["TT_LEARN_STOP"] = "0"
for k, name in ((1, "re-pack stage A (Gram, W2 column statistics, fp32 images)"), (2, "re-pack stage C (Cholesky, operand images)")):
    os.environ["TT_PACK_STOP"] = str(k)
    ln.learn(repack_into="agent")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ln.learn(repack_into="agent")
    t = timed(g.replay)
    print(f"{14 + k:2d} {name:58s} cumulative {t:7.1f} us   this launch {t - prev:6.1f} us")
    prev = t
