"""Actor forward timing (CUDA events, N = 2^22) for every precision mode / kernel variant."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddpg_trucktrailer_b200 as tt
_n = [a for a in sys.argv[1:] if not a.startswith('--')]
N = int(_n[0]) if _n else 1 << 22
actor = tt.agent.CudaActor(); actor.load_state_dict(tt.init_actor_state_dict(seed=0))
obs = torch.empty(N, 23, device="cuda").uniform_(-1, 1)
out = torch.empty(N, device="cuda")
store = "--store" in sys.argv          # also time the fused replay store of s (what tt_rollout_step launches)
if store:
    import ctypes as C
    from ddpg_trucktrailer_b200 import _lib
    L = tt.load()
    cap = 1 << 24
    S = torch.empty(cap, 23, device="cuda")
    ring = _lib.ReplayRing(S.data_ptr(), 0, 0, 0, 0, cap, 0)
    def fwd(prec):
        _lib.check(L.tt_actor_forward_store(actor._h, obs.data_ptr(), 23, N, out.data_ptr(), _lib.PRECISIONS[prec], C.byref(ring), _lib.stream_ptr()))
else:
    def fwd(prec):
        actor.forward(obs, out=out, precision=prec)
for prec in ("f16", "f16_plain", "bf16") + (("fp32",) if "--fp32" in sys.argv else ()):
    sustained = "--sustained" in sys.argv     # 300 launches first: the GPU reaches its power cap (what a long rollout sees)
    for _ in range(300 if sustained else 3): fwd(prec)
    if not sustained: torch.cuda.synchronize()
    ts = []
    if sustained:
        evs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fwd(prec)
            e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) / 10 for a, b in evs]
    else:
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fwd(prec); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"sustained={sustained} store={store} TT_TC_VARIANT={os.environ.get('TT_TC_VARIANT','default')} {prec}: {ms:.3f} ms  {N/ms/1e6:.2f} Grows/s  {N*259000/ms/1e9:.0f} TFLOP/s  all={['%.3f'%t for t in ts]}")
