"""BASELINE.json configs[4]: env-step / rollout throughput sweep over N = 2^12 .. 2^24 envs per GPU (CUDA events).
Writes a markdown table to stdout.  usage: python profiles/sweep.py [max_log2]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import _lib

hi = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L = tt.load()
print("| envs/GPU | env step only (us) | env Gsteps/s | env % HBM roofline (229 B) | rollout f16 (us/iter) | rollout Msteps/s | same as one CUDA graph of K iterations (us/iter) | Msteps/s | rollout fp32-actor (us/iter) | Msteps/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for lg in range(12, hi + 1, 2):
    N = 1 << lg
    res = {}
    for prec in ("f16", "fp32"):
        if prec == "fp32" and lg > 20:
            res[prec] = float("nan"); continue
        env = tt.VecTruckTrailerEnv(N, seed=27)
        ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=4 * N, actor_seed=0, precision=prec)
        eng = tt.RolloutEngine(env, ag); eng.reset()
        iters = 200 if lg <= 16 else (50 if lg <= 20 else 15)
        for _ in range(5): eng.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): eng.step()
        e1.record(); torch.cuda.synchronize()
        res[prec] = e0.elapsed_time(e1) / iters * 1e3
        if prec == "f16":
            k = eng.capture()
            reps = max(iters // k, 3)
            eng.step_graph(); torch.cuda.synchronize()
            e0.record()
            for _ in range(reps): eng.step_graph()
            e1.record(); torch.cuda.synchronize()
            res["graph"] = e0.elapsed_time(e1) / (reps * k) * 1e3
            s = _lib.stream_ptr()
            tot = 0.0
            for _ in range(iters):
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                _lib.check(L.tt_env_step(env._h, eng.scaled.data_ptr(), env._obs[0].data_ptr(), 23, env._reward.data_ptr(), env._done.data_ptr(), None, s))
                a1.record(); torch.cuda.synchronize()
                tot += a0.elapsed_time(a1)
                _lib.check(L.tt_env_reset(env._h, env._done.data_ptr(), env._obs[0].data_ptr(), 23, s)); env.tick()   # (C-level masked reset does not tick)
            env_us = tot / iters * 1e3
        del eng, ag, env
        torch.cuda.empty_cache()
    print(f"| 2^{lg} = {N} | {env_us:.1f} | {N/env_us/1e3:.2f} | {N*229/env_us/1e3/6553.3*100:.1f} | {res['f16']:.1f} | {N/res['f16']:.1f} | {res['graph']:.1f} | {N/res['graph']:.1f} | {res['fp32']:.1f} | {N/res['fp32']:.1f} |", flush=True)
