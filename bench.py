#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 truck-trailer rollout path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU implementation of the same path

Metric (BASELINE.json): env-steps/sec of the fused env+actor rollout.  One "step" = one rollout iteration
(actor forward -> OU noise -> clip*pi/4 -> env step -> replay store -> reset of finished envs, i.e. the body of
DDPG/trainv2.py:511-531 without learn()) over all environments of a rank; value = environments * K / time,
summed over ranks (weak scaling: 2^22 environments per GPU, sharded by global env id, no data-path collective).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL's own messages (the "NCCL version ..." banner is printed from NCCL_DEBUG=VERSION
# upwards, and the image sets it) go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":          # the banner itself ignores NCCL_DEBUG_FILE
    del os.environ["NCCL_DEBUG"]

def _baseline_metric():
    """BASELINE.json's headline metric string, verbatim (the primary quantity of it: env-steps/sec of the fused rollout)."""
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "env-steps/sec (fused env+actor) at 1/2/4/8 B200; % HBM roofline"


METRIC = _baseline_metric()
UNIT = "env-steps/s"
# algorithmic bytes / flops per env-step (restated in DESIGN.md)
ENV_BYTES, OU_BYTES, ACTOR_BYTES, STORE_BYTES = 229, 16, 96, 386
ACTOR_FLOPS = 259_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 22, help="environments per GPU")
    ap.add_argument("--precision", default=os.environ.get("TT_BENCH_PRECISION", "f16"), choices=["fp32", "bf16", "f16"])
    ap.add_argument("--ring", type=int, default=1 << 24, help="replay ring capacity per GPU (transitions)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--config4", action="store_true", help="also time BASELINE.json configs[3]: rollout + replay store + one DDPG "
                    "critic/actor update (batch 64) per iteration + stats all-reduce + actor broadcast (extra key `config4`)")
    ap.add_argument("--preroll", type=int, default=256, help="untimed rollout iterations before the warm-up, so that episodes "
                    "terminate and reset at their steady-state rate inside the timed region")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1, load0=None, load1=None):
        """Clocks inside the timed region [t0, t1] (the region lasts tens of ms, nvidia-smi samples every 20 ms); if no
        sample fell into it, the samples taken under the same load right around it (pre-roll .. end of the e2e run)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.02 <= t <= t1 + 0.02]
        if not rows and load0 is not None:
            rows = [r for t, r in self.rows if load0 <= t <= load1]
        rows = rows or [r for _, r in self.rows[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU arm ----
def cpu_port(n_envs, seconds, threads, steps=None, warmup=1):
    """The oracle's CPU port of the whole rollout iteration (kind "port": the reference is pure Python and is
    not present on the GPU box), all host threads, on a bounded sample of the workload."""
    import numpy as np
    from oracle import oracle as orc
    rng = np.random.default_rng(0)            # reference init distributions (networks.py:110-131)
    u = lambda shape, f: rng.uniform(-f, f, shape).astype(np.float32)
    sd = {"fc1.weight": u((400, 23), 0.05), "fc1.bias": u(400, 0.05), "bn1.weight": np.ones(400, np.float32),
          "bn1.bias": np.zeros(400, np.float32), "fc2.weight": u((300, 400), 300 ** -0.5), "fc2.bias": u(300, 300 ** -0.5),
          "bn2.weight": np.ones(300, np.float32), "bn2.bias": np.zeros(300, np.float32), "mu.weight": u((1, 300), 0.003),
          "mu.bias": u(1, 0.003)}
    port = orc.RolloutPort(n_envs, sd, seed=27, threads=threads, capacity=max(n_envs, 1 << 16))
    for _ in range(warmup):
        port.step()
    t0, it = time.perf_counter(), 0
    while True:
        port.step(); it += 1
        if (steps is not None and it >= steps) or (steps is None and time.perf_counter() - t0 >= seconds):
            break
    dt = time.perf_counter() - t0
    return n_envs * it / dt, it, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 1 << 15                                   # bounded sample: 32768 envs per step
    t0 = time.perf_counter()
    val, it, dt = cpu_port(n, 0, threads, steps=args.steps, warmup=max(args.warmup, 1))
    sample = f"{n} envs x {it} rollout iterations (actor fp32 + OU + env step float64 RK45 + store + reset) in {dt:.1f}s"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / it, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "full rollout: actor(23-400-300-1)+OU+simv2 step+reward_functionv1+replay store+reset, "
                                   "2^22 envs/GPU (CPU arm: bounded sample)", "envs_per_step_sample": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm ----
def run_b200(args):
    import torch
    import torch.distributed as dist
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import _lib
    from ddpg_trucktrailer_b200 import dist as ttd

    rank, world, local = ttd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.envs
    offset = rank * N
    L = tt.load()
    pk = peaks()

    env = tt.VecTruckTrailerEnv(N, seed=27, global_env_offset=offset, device=dev)
    agent = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=args.ring, num_envs=N, device=dev, seed=27, global_env_offset=offset,
                        precision=args.precision, actor_seed=0)
    sd = tt.init_actor_state_dict(seed=0)
    if world > 1:   # NCCL over NVLink: the only collectives of the path (actor broadcast, stats all-reduce)
        sd = ttd.broadcast_actor({k: v.to(dev) for k, v in sd.items()}, src=0, device=dev)
    agent.load_actor_state_dict(sd)
    eng = tt.RolloutEngine(env, agent, store=True)
    eng.reset()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, iters):
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(); barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    W, K = max(args.warmup, 3), args.steps
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t_load0 = time.time()
    for _ in range(args.preroll):          # population reaches its stationary mix of episode ages (max episode ~ 250 steps)
        eng.step()
    env.stats_tensor(clear=True)
    for _ in range(W):
        eng.step()
    torch.cuda.synchronize()

    # ---- headline: K rollout iterations, everything resident in HBM ----
    launches0 = L.tt_launch_count()
    t_wall0 = time.time()
    ms = timed(eng.step, K)
    t_wall1 = time.time()
    launches = L.tt_launch_count() - launches0
    value = world * N * K / (ms * 1e-3)
    stats = ttd.all_reduce_stats(env.stats_tensor(clear=True).clone())

    # ---- e2e: the same iteration driven from the host with HOST buffers inside the timed region:
    #      H2D of the step's external input (the current actor parameters from pinned host memory, re-packed on the
    #      device) and D2H of the step's results the reference driver reads (reward and done of every env + stats)
    flat_host = ttd.flatten_actor(sd).cpu().pin_memory()
    # the policy of step t + 1 is uploaded and re-packed on a side stream WHILE step t runs (two packed actors, used
    # alternately): what an asynchronous learner hands over; every byte still moves, every step, inside the timed region
    flat_dev = [torch.empty_like(flat_host, device=dev) for _ in range(2)]
    views = [ttd.unflatten_actor(f, sd) for f in flat_dev]
    actors = [agent.actor, tt.agent.CudaActor(*agent.actor.dims, device=dev)]
    up_stream = torch.cuda.Stream(device=dev)
    packed = [torch.cuda.Event() for _ in range(2)]
    step_done = torch.cuda.Event()

    def upload(slot):                                                  # H2D + device re-pack of one policy, on up_stream
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(step_done)                            # the previous user of this slot's images has finished
            flat_dev[slot].copy_(flat_host, non_blocking=True)
            actors[slot].load_state_dict(views[slot])
            packed[slot].record(up_stream)
    # results are read back one iteration behind on a copy stream (double-buffered staging), so the PCIe transfer of
    # step t overlaps the kernels of step t+1; every byte still moves inside the timed region
    rew_host = [torch.empty(N, dtype=torch.float32).pin_memory() for _ in range(2)]
    done_host = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(2)]
    stats_host = [torch.empty(16, dtype=torch.float64).pin_memory() for _ in range(2)]
    rew_stage = [torch.empty(N, dtype=torch.float32, device=dev) for _ in range(2)]
    done_stage = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(2)]
    stats_stage = [torch.empty(16, dtype=torch.float64, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [torch.cuda.Event() for _ in range(2)]
    drained = [torch.cuda.Event() for _ in range(2)]
    for ev in drained:
        ev.record()
    e2e_it = [0]

    def e2e_step():
        b = e2e_it[0] & 1
        e2e_it[0] += 1
        main = torch.cuda.current_stream()
        main.wait_event(packed[b])                                   # this step's policy: uploaded + re-packed during the previous step
        agent.actor = actors[b]
        step_done.record(main)                                       # (everything before this step, incl. the last user of slot b ^ 1)
        upload(b ^ 1)                                                # H2D + re-pack of the NEXT step's policy, overlapped with this step
        _, r, d = eng.step()
        main.wait_event(drained[b])                                  # staging buffer b was read out two steps ago
        rew_stage[b].copy_(r); done_stage[b].copy_(d); stats_stage[b].copy_(env.stats_tensor(clear=True))
        staged[b].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged[b])
            rew_host[b].copy_(rew_stage[b], non_blocking=True)       # D2H: what the reference driver reads each step
            done_host[b].copy_(done_stage[b], non_blocking=True)
            stats_host[b].copy_(stats_stage[b], non_blocking=True)
            drained[b].record(copy_stream)

    def e2e_run(iters):
        for _ in range(iters):
            e2e_step()
        copy_stream.synchronize(); up_stream.synchronize()

    step_done.record()
    upload(0)                                                          # the first step's policy

    e2e_run(2)
    torch.cuda.synchronize()
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    e2e_run(K)
    torch.cuda.current_stream().wait_stream(copy_stream)
    torch.cuda.current_stream().wait_stream(up_stream)               # every upload issued inside the region also ends inside it
    ev1.record()
    torch.cuda.synchronize(); barrier()
    t_e2e = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e)
    agent.actor = actors[0]
    clocks = sampler.stop(t_wall0, t_wall1, t_load0, time.time()) if sampler else None     # sampled from the pre-roll to the end of e2e
    e2e = {"value": world * N * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": flat_host.numel() * 4,
           "d2h_bytes_per_step": N * 5 + 128, "ms_per_step": ms_e2e / K,
           "note": "host buffers: actor parameters H2D + re-pack every step (side stream, one step ahead, two packed actors); reward/done/stats D2H every step (copy stream, one step behind)"}

    # ---- per-kernel timing (CUDA events on the launching stream) -> roofline of the dominant kernel ----
    s = _lib.stream_ptr()
    cur, nxt = env._obs[env._cur], env._obs[env._cur ^ 1]
    m = agent.memory
    prec = _lib.PRECISIONS[args.precision]
    import ctypes as C
    ring = _lib.ReplayRing(m.state_memory.data_ptr(), m.action_memory.data_ptr(), m.reward_memory.data_ptr(),
                           m.new_state_memory.data_ptr(), m.terminal_memory.data_ptr(), m.mem_size, m.mem_cntr)
    rp = C.byref(ring)
    # the two kernels of one tt_rollout_step, each timed alone (the replay store, the OU noise, the reset of finished envs
    # and the iteration tick are fused into them)
    kern = {
        "actor": lambda: _lib.check(L.tt_actor_choose_action(agent.actor._h, cur.data_ptr(), env.ld_obs, N, agent.noise.x_prev.data_ptr(), 27, offset,
                                                             agent.noise.iter_ptr, 0, eng.action.data_ptr(), eng.scaled.data_ptr(), prec, rp, s)),
        "env_step": lambda: _lib.check(L.tt_env_step_reset(env._h, eng.scaled.data_ptr(), nxt.data_ptr(), env.ld_obs, env._reward.data_ptr(),
                                                           env._done.data_ptr(), agent.noise.x_prev.data_ptr(), rp, s)),
        # for reference only (NOT part of the fused rollout): the stand-alone ring scatter kernel
        "replay_store_standalone": lambda: _lib.check(L.tt_replay_store(m.state_memory.data_ptr(), m.action_memory.data_ptr(), m.reward_memory.data_ptr(),
                                                             m.new_state_memory.data_ptr(), m.terminal_memory.data_ptr(), m.mem_size, m.mem_cntr,
                                                             cur.data_ptr(), env.ld_obs, eng.action.data_ptr(), env._reward.data_ptr(),
                                                             nxt.data_ptr(), env.ld_obs, env._done.data_ptr(), N, s)),
    }
    # Every kernel is timed IN the running rollout (same power / clock state as the headline: under sustained load a B200
    # settles well below its boost clock): one untimed rollout iteration, then the kernel once more between two events,
    # K times back to back without host synchronisation.  (Timing a kernel alone in a cold 30 ms burst flatters it by 10 %.)
    kms = {}
    for name, fn in kern.items():
        for _ in range(60):
            eng.step()
        evs = []
        for _ in range(K):
            eng.step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        kms[name] = sum(a.elapsed_time(b) for a, b in evs) / K
    # algorithmic work per launch: env step 229 B + its share of the fused store (s', r, done: 97 B)
    algo = {"actor": ("tensor", N * ACTOR_FLOPS / 1e12), "env_step": ("hbm", N * (ENV_BYTES + 97) / 1e9),
            "replay_store_standalone": ("hbm", N * STORE_BYTES / 1e9)}
    kernels = {}
    for name, (bound, work) in algo.items():
        peak = pk["hbm"] if bound == "hbm" else pk["tf_sust"]     # kernels timed inside the running rollout: sustained figure
        ach = work / (kms[name] * 1e-3)
        kernels[name] = {"ms": kms[name], "bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                         "frac": ach / peak}
    in_step = ["actor", "env_step"]
    dom = max([k for k in algo if k in in_step], key=lambda n: kms[n])
    try:      # DRAM bytes per env from the committed `ncu --set full` capture (profiles/), scaled to this launch
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)["bytes_per_env"].get(dom)
        traffic = None if traffic is None else traffic * N
    except Exception:
        traffic = None
    roofline = {"kernel": dom, "bound": kernels[dom]["bound"], "achieved": kernels[dom]["achieved"], "peak": kernels[dom]["peak"],
                "unit": kernels[dom]["unit"], "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": pk["src"],
                "share_of_step": kms[dom] / sum(kms[k] for k in in_step)}

    # ---- optional: BASELINE.json configs[3] = the rollout iteration + one learner update (batch 64) per iteration ----
    config4 = None
    if args.config4:
        agent.learner_graph = True                   # learner.py: the whole DDPG update as one CUDA graph, samples rank-local ring
        for _ in range(3):
            agent.learn()
        torch.cuda.synchronize()
        stats_dev = torch.zeros(16, dtype=torch.float64, device=dev)

        def c4_step():
            eng.step()
            agent.learn()                                              # replays the graph (incl. the actor re-pack) on the same stream
            if world > 1:                                              # section 8e: the only collectives of the path
                stats_dev.copy_(env.stats_tensor(clear=True)); dist.all_reduce(stats_dev)
                flat = ttd.flatten_actor(agent.actor.state_dict(), device=dev); dist.broadcast(flat, src=0)
                agent.load_actor_state_dict(ttd.unflatten_actor(flat, sd))
        for _ in range(3):
            c4_step()
        ms4 = timed(c4_step, K)
        config4 = {"value": world * N * K / (ms4 * 1e-3), "unit": UNIT, "ms_per_step": ms4 / K,
                   "note": "configs[3]: rollout + fused replay store + one DDPG update (batch 64, learner step as one CUDA graph, "
                           "sequential on the rollout stream)" + (" + stats all-reduce + actor broadcast (NCCL)" if world > 1 else "")}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, it, dt = cpu_port(1 << 14, args.cpu_seconds, threads)
        v1, it1, dt1 = cpu_port(1 << 12, min(args.cpu_seconds, 6.0), 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{1 << 14} envs x {it} rollout iterations in {dt:.1f}s on {threads} threads (oracle C port: actor fp32 + OU + "
                         f"float64 adaptive-RK45 env step + store + reset)",
               "single_thread_value": v1}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16": "bf16", "f16": "f16"}[args.precision],
                "data": "synthetic",
                "config": {"workload": f"full rollout: actor(23-400-300-1,{args.precision})+OU+simv2 step(f64/f32 DP5)+reward_functionv1+"
                                       f"replay store (fused into the producers)+auto-reset, {N} envs/GPU", "envs_per_gpu": N, "ring_capacity": args.ring, "preroll_iterations": args.preroll,
                           "l2": "working set per step (>1.2 GB/GPU) far exceeds the 126 MB L2; no flush needed",
                           "sharding": "global env id ranges, no data-path collective; NCCL only for actor broadcast + stats all-reduce"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
                "cpu_baseline": cpu, "rollout_stats": ttd.summarize(stats)}
        if config4 is not None:
            line["config4"] = config4
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
