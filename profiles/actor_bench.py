"""Actor forward timing (CUDA events, N = 2^22) for every precision mode / kernel variant."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddpg_trucktrailer_b200 as tt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
actor = tt.agent.CudaActor(); actor.load_state_dict(tt.init_actor_state_dict(seed=0))
obs = torch.empty(N, 23, device="cuda").uniform_(-1, 1)
out = torch.empty(N, device="cuda")
for prec in ("f16", "bf16") + (("fp32",) if "--fp32" in sys.argv else ()):
    for _ in range(3): actor.forward(obs, out=out, precision=prec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); actor.forward(obs, out=out, precision=prec); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"TT_TC_VARIANT={os.environ.get('TT_TC_VARIANT','default')} {prec}: {ms:.3f} ms  {N/ms/1e6:.2f} Grows/s  {N*259000/ms/1e9:.0f} TFLOP/s  all={['%.3f'%t for t in ts]}")
