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
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="bounded sample of the oracle C port")
    ap.add_argument("--ref-steps", type=int, default=20000, help="env steps of the single-process Python reference loop (BASELINE.md section 3)")
    ap.add_argument("--no-config4", action="store_true", help="skip BASELINE.json configs[3] (rollout + replay store + one DDPG critic/actor "
                    "update (batch 64) per iteration + stats all-reduce + actor broadcast; key `config4`)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the mini-sweep over envs per GPU (BASELINE.json configs[4]; key `sweep`)")
    ap.add_argument("--e2e-preroll-s", type=float, default=0.5, help="seconds of the e2e pipeline itself run right before its timed region")
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

    def window(self, t0, t1):
        """Clocks of the samples taken inside [t0, t1] (no fallback); None if there were none."""
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        if not rows:
            return None
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}

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


# ------------------------------------------------------------------------------------------------ CPU arms ----
WORKLOAD = ("full rollout: actor(23-400-300-1)+OU noise+clip*pi/4+simv2 step+reward_functionv1+replay store+reset of finished "
            "episodes (DDPG/trainv2.py:489-531 without learn()), 2^22 envs/GPU")


def _actor_sd_numpy():
    import numpy as np
    rng = np.random.default_rng(0)            # reference init distributions (networks.py:110-131)
    u = lambda shape, f: rng.uniform(-f, f, shape).astype(np.float32)
    return {"fc1.weight": u((400, 23), 0.05), "fc1.bias": u(400, 0.05), "bn1.weight": np.ones(400, np.float32),
            "bn1.bias": np.zeros(400, np.float32), "fc2.weight": u((300, 400), 300 ** -0.5), "fc2.bias": u(300, 300 ** -0.5),
            "bn2.weight": np.ones(300, np.float32), "bn2.bias": np.zeros(300, np.float32), "mu.weight": u((1, 300), 0.003),
            "mu.bias": u(1, 0.003)}


def cpu_port(n_envs, seconds, threads, steps=None, warmup=1):
    """The oracle's C port of the whole rollout iteration (kind "port"), all host threads, on a bounded sample."""
    from oracle import oracle as orc
    port = orc.RolloutPort(n_envs, _actor_sd_numpy(), seed=27, threads=threads, capacity=max(n_envs, 1 << 16))
    for _ in range(warmup):
        port.step()
    t0, it = time.perf_counter(), 0
    while True:
        port.step(); it += 1
        if (steps is not None and it >= steps) or (steps is None and time.perf_counter() - t0 >= seconds):
            break
    dt = time.perf_counter() - t0
    return n_envs * it / dt, it, dt


def python_reference_available():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import prepare_ref
    return prepare_ref.available()


def python_reference(mode, procs, steps_per_proc, warm=300, chunks=1):
    """The UNMODIFIED reference Python loop (baseline/_ref, baseline/ref_loop.py): `procs` processes, OMP threads = 1."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loop
    return ref_loop.time_loop(mode, procs, steps_per_proc, warm=warm, chunks=chunks)


def cpu_baseline_block(args):
    """cpu_baseline of the b200 arm (rank 0, N = 1): the reference's own Python loop timed on this box's host cores in the same
    run -- single process and one process per core, full rollout loop and env-only loop -- with the
    oracle's C port of the same iteration as a second figure."""
    cores = os.cpu_count() or 1
    out = {}
    vp, itp, dtp = cpu_port(1 << 14, args.cpu_seconds, cores)
    vp1, _, _ = cpu_port(1 << 12, min(args.cpu_seconds, 4.0), 1)
    port = {"value": vp, "unit": UNIT, "cores": cores, "kind": "port", "single_thread_value": vp1,
            "sample": f"{1 << 14} envs x {itp} rollout iterations in {dtp:.1f}s on {cores} threads (oracle C port: actor fp32 + OU + float64 "
                      f"adaptive-RK45 env step + store + reset)"}
    if not python_reference_available():
        port["note"] = "baseline/_ref is missing (run __graft_entry__.build() where /root/reference is mounted): the Python reference was not timed"
        return port
    single = python_reference("rollout", 1, args.ref_steps, warm=1000)
    per_proc = max(2000, args.ref_steps // 4)
    multi = python_reference("rollout", cores, per_proc, warm=300)
    env1 = python_reference("env", 1, args.ref_steps, warm=1000)
    envm = python_reference("env", cores, per_proc, warm=300)
    out = {"value": multi["value"], "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": f"UNMODIFIED reference Python loop (choose_action -> clip*high -> env.step -> remember, reset on done; actor on the CPU): "
                     f"{cores} processes (one per core, OMP threads 1) x {per_proc} env steps in {multi['seconds']:.1f}s",
           "single_process_value": single["value"],
           "single_process_sample": f"1 process x {single['steps']} env steps in {single['seconds']:.1f}s",
           "env_only": {"single_process_value": env1["value"], "all_cores_value": envm["value"],
                        "sample": f"env.step with U(-pi/4, pi/4) steering + reset on done: 1 x {env1['steps']} steps in {env1['seconds']:.1f}s; "
                                  f"{cores} x {per_proc} steps in {envm['seconds']:.1f}s"},
           "port": port}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.  One bench "step" = every
    process advances its own environment by a bounded chunk of env steps of the unmodified Python rollout loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    W, K = max(args.warmup, 1), args.steps
    if python_reference_available():
        chunk = 400                                  # env steps per process and bench step (~0.4 s of the ~1 k steps/s loop)
        res = python_reference("rollout", cores, chunk * K, warm=chunk * W, chunks=K)
        val, secs = res["value"], res["seconds"]
        kind, dtype = "reference", "f64"
        sample = (f"UNMODIFIED reference Python loop (baseline/_ref: DDPG_agent.choose_action -> clip*high -> simv2 env.step -> remember, "
                  f"reset on done; actor on the CPU): {cores} processes (one per core, OMP threads 1) x {K} steps x {chunk} env steps "
                  f"after {W} warm-up steps, {secs:.1f}s")
        extra = {"envs_per_step_sample": cores * chunk}
    else:
        n = 1 << 15
        val, it, secs = cpu_port(n, 0, cores, steps=K, warmup=W)
        kind, dtype = "port", "f64"
        sample = f"baseline/_ref missing -> oracle C port: {n} envs x {it} rollout iterations in {secs:.1f}s"
        extra = {"envs_per_step_sample": n}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": dict({"workload": WORKLOAD + " (CPU arm: bounded sample of the same loop)"}, **extra),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm ----
def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and thereby its pinned staging buffers, by first touch) to the CPUs of the NUMA node its
    GPU hangs off -- as far as the process is allowed to: a container cpuset that covers one node only cannot be left.  Returns
    what was found / done (reported under e2e.host)."""
    info = {"gpu_numa_node": None, "allowed_cpus": len(os.sched_getaffinity(0)), "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        sysdir = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        with open(sysdir + "/numa_node") as f:
            info["gpu_numa_node"] = int(f.read())
        with open(sysdir + "/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        ok = cpus & os.sched_getaffinity(0)
        info["gpu_local_cpus_allowed"] = len(ok)
        if ok and ok != os.sched_getaffinity(0):
            os.sched_setaffinity(0, ok)
            info["bound"] = True
    except Exception as e:                            # no NVML / sysfs entry: nothing to bind to
        info["note"] = f"{type(e).__name__}: {e}"[:120]
    return info


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import _lib
    from ddpg_trucktrailer_b200 import dist as ttd

    rank, world, local = ttd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # before any pinned allocation (first touch decides where the pages live); single-GPU runs keep the whole host for the CPU baseline
    numa = bind_to_gpu_numa_node(local) if int(os.environ.get("WORLD_SIZE", "1")) > 1 else {"bound": False, "allowed_cpus": len(os.sched_getaffinity(0))}
    N = args.envs
    offset = rank * N
    L = tt.load()
    pk = peaks()
    W, K = max(args.warmup, 3), args.steps

    env = tt.VecTruckTrailerEnv(N, seed=27, global_env_offset=offset, device=dev)
    agent = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=args.ring, num_envs=N, device=dev, seed=27, global_env_offset=offset,
                        precision=args.precision, actor_seed=0, allow_out_of_bar=args.precision == "bf16")
    sd = tt.init_actor_state_dict(seed=0)
    if world > 1:   # NCCL over NVLink: the only collectives of the path (actor broadcast, stats all-reduce)
        sd = ttd.broadcast_actor({k: v.to(dev) for k, v in sd.items()}, src=0, device=dev)
    agent.load_actor_state_dict(sd)
    eng = tt.RolloutEngine(env, agent, store=True)
    eng.reset()
    sync = ttd.OverlappedSync(dev, src=0)            # side-stream collectives (no-ops at world == 1)

    # ---- everything the later phases need is allocated NOW, so that no allocation sits between a pre-roll and its timed region
    flat_host = ttd.flatten_actor(sd).cpu().pin_memory()
    flat_dev = [torch.empty_like(flat_host, device=dev) for _ in range(2)]
    actors = [agent.actor, tt.agent.CudaActor(*agent.actor.dims, device=dev)]
    rew_host = [torch.empty(N, dtype=torch.float32).pin_memory() for _ in range(2)]
    nbits = (N + 31) // 32                                            # done is read back bit-packed (tt_env_set_done_bits): 1 bit per env over PCIe
    done_host = [torch.empty(nbits, dtype=torch.int32).pin_memory() for _ in range(2)]
    stats_host = [torch.empty(16, dtype=torch.float64).pin_memory() for _ in range(2)]
    rew_stage = [torch.empty(N, dtype=torch.float32, device=dev) for _ in range(2)]
    done_stage = [torch.zeros(nbits, dtype=torch.int32, device=dev) for _ in range(2)]
    stats_stage = [torch.empty(16, dtype=torch.float64, device=dev) for _ in range(2)]
    up_stream, copy_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, iters):
        """EXACTLY `iters` steps between a barrier + synchronize on both sides, CUDA events, max over ranks; also the wall-clock
        window of the region (for the clock samples)."""
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(); w1 = time.time(); barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), (w0, w1)

    def headline_step():
        eng.step()
        if world > 1:      # section 8e: the per-iteration statistics all-reduce, issued on the side stream, consumed one iteration behind
            sync.push_stats(env.stats_tensor(clear=True))

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t_load0 = time.time()
    for _ in range(args.preroll):          # population reaches its stationary mix of episode ages (max episode ~ 250 steps)
        headline_step()
    stats_acc = torch.zeros(16, dtype=torch.float64, device=dev)
    env.stats_tensor(clear=True)
    for _ in range(W):
        headline_step()

    # ---- headline: K rollout iterations, everything resident in HBM ----
    launches0 = L.tt_launch_count()
    ms, win_value = timed(headline_step, K)
    launches = L.tt_launch_count() - launches0
    value = world * N * K / (ms * 1e-3)
    if world > 1:
        stats = sync.flush_stats().clone()
    else:
        stats = env.stats_tensor(clear=True).clone()

    # ---- e2e: the same iteration driven from the host with HOST buffers inside the timed region:
    #      H2D of the step's external input (the current actor parameters from pinned host memory, re-packed on the
    #      device) and D2H of the step's results the reference driver reads (reward and done -- bit-packed -- of every env + stats).
    # The policy of step t + 1 is uploaded and re-packed on a side stream WHILE step t runs (two packed actors, used
    # alternately): what an asynchronous learner hands over; results are read back one iteration behind on a copy stream
    # (two pairs of result buffers, written by the kernels alternately).  Every byte still moves, every step, inside the timed region.
    packed = [torch.cuda.Event() for _ in range(2)]
    step_done = torch.cuda.Event()
    staged = [torch.cuda.Event() for _ in range(2)]
    drained = [torch.cuda.Event() for _ in range(2)]
    for ev in drained:
        ev.record()
    e2e_it = [0]

    def upload(slot):                                                  # H2D + device re-pack of one policy, on up_stream
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(step_done)                            # the previous user of this slot's images has finished
            flat_dev[slot].copy_(flat_host, non_blocking=True)
            actors[slot].load_flat(flat_dev[slot])
            packed[slot].record(up_stream)

    def e2e_step():
        b = e2e_it[0] & 1
        e2e_it[0] += 1
        main = torch.cuda.current_stream()
        main.wait_event(packed[b])                                   # this step's policy: uploaded + re-packed during the previous step
        agent.actor = actors[b]
        step_done.record(main)                                       # (everything before this step, incl. the last user of slot b ^ 1)
        upload(b ^ 1)                                                # H2D + re-pack of the NEXT step's policy, overlapped with this step
        main.wait_event(drained[b])                                  # result buffers b were read out two steps ago
        eng.step(out=(rew_stage[b], None, done_stage[b]))            # the kernels write this step's reward / bit-packed done straight into pair b
        stats_stage[b].copy_(env.stats_tensor(clear=True))
        staged[b].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged[b])
            rew_host[b].copy_(rew_stage[b], non_blocking=True)       # D2H: what the reference driver reads each step
            done_host[b].copy_(done_stage[b], non_blocking=True)
            stats_host[b].copy_(stats_stage[b], non_blocking=True)
            drained[b].record(copy_stream)

    # what bounds e2e at 8 GPUs is the host side (17 MB D2H per rank and step through one host): measure the D2H rate every rank
    # gets while ALL ranks copy at once
    d2h_bytes = N * 4 + nbits * 4                                     # reward + bit-packed done per step (+ 128 B of statistics)
    d2h_probe = None
    if world > 1:
        barrier(); torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(20):
            rew_host[0].copy_(rew_stage[0], non_blocking=True); done_host[0].copy_(done_stage[0], non_blocking=True)
        p1.record(); torch.cuda.synchronize()
        gbs = torch.tensor([20 * d2h_bytes / (p0.elapsed_time(p1) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        lo, hi = gbs.clone(), gbs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(gbs)
        d2h_probe = {"per_rank_min": float(lo), "per_rank_max": float(hi), "aggregate": float(gbs),
                     "needed_per_rank_at_value": d2h_bytes / (ms / K * 1e-3) / 1e9}
    step_done.record()
    upload(0)                                                          # the first step's policy
    # its own pre-roll: the e2e pipeline itself for >= args.e2e_preroll_s, straight into the timed region (same clock / power
    # state as the region; the only gap is the contract's barrier + synchronize)
    pre = max(W, int(args.e2e_preroll_s / max(ms / K * 1e-3, 1e-6)))
    for _ in range(pre):
        e2e_step()
    barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w_e0 = time.time()
    ev0.record()
    for _ in range(K):
        e2e_step()
    torch.cuda.current_stream().wait_stream(copy_stream)
    torch.cuda.current_stream().wait_stream(up_stream)               # every copy issued inside the region also ends inside it
    ev1.record()
    torch.cuda.synchronize(); w_e1 = time.time(); barrier()
    t_e2e = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e)
    agent.actor = actors[0]
    actors[0].load_state_dict(sd)
    env.set_done_bits(None)
    env.stats_tensor(clear=True)
    clocks = clocks_e2e = None
    if sampler:
        # the regions last tens of ms and nvidia-smi samples every 20 ms: take the samples of the region plus the 0.3 s of the
        # same uninterrupted load right before it
        clocks = sampler.window(win_value[0] - 0.3, win_value[1]) or sampler.window(t_load0, win_value[1])
        clocks_e2e = sampler.window(w_e0 - 0.3, w_e1)
    e2e = {"value": world * N * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": flat_host.numel() * 4,
           "d2h_bytes_per_step": d2h_bytes + 128, "ms_per_step": ms_e2e / K, "preroll_iterations": pre, "clocks": clocks_e2e,
           "host": dict(numa, d2h_gbs_all_ranks_copying=d2h_probe),
           "note": "host buffers: actor parameters H2D + re-pack every step (side stream, one step ahead, two packed actors); reward / bit-packed done / stats D2H "
                   "every step (copy stream, one step behind); timed right after its own pre-roll of the same pipeline"}

    # ---- per-kernel timing (CUDA events on the launching stream) -> roofline of the dominant kernel ----
    s = _lib.stream_ptr()
    cur, nxt = env._obs[env._cur], env._obs[env._cur ^ 1]
    m = agent.memory
    prec = _lib.PRECISIONS[args.precision]
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
    # algorithmic work per launch: actor 259 000 FLOP / row; env step 229 B + its share of the fused store
    # (s', r, done: 97 B); the noise state (8 B), the stored action / scaled action and the reset are NOT counted
    algo = {"actor": ("tensor", N * ACTOR_FLOPS / 1e12), "env_step": ("hbm", N * (ENV_BYTES + 97) / 1e9)}
    kernels = {}
    for name, (bound, work) in algo.items():
        peak = pk["hbm"] if bound == "hbm" else pk["tf_sust"]     # kernels timed inside the running rollout: sustained figure
        ach = work / (kms[name] * 1e-3)
        kernels[name] = {"ms": kms[name], "bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                         "frac": ach / peak}
    kernels["env_step"]["frac_of_bare_step_229B"] = N * ENV_BYTES / 1e9 / (kms["env_step"] * 1e-3) / pk["hbm"]
    in_step = ["actor", "env_step"]
    dom = max(in_step, key=lambda n: kms[n])
    try:      # DRAM bytes per env from the committed `ncu --set full` capture (profiles/), scaled to this launch
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)["bytes_per_env"].get(dom)
        traffic = None if traffic is None else traffic * N
    except Exception:
        traffic = None
    roofline = {"kernel": dom, "bound": kernels[dom]["bound"], "achieved": kernels[dom]["achieved"], "peak": kernels[dom]["peak"],
                "unit": kernels[dom]["unit"], "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": pk["src"],
                "share_of_step": kms[dom] / sum(kms[k] for k in in_step)}

    # ---- BASELINE.json configs[3]: the rollout iteration + one DDPG update (batch 64) per iteration (+ collectives) ----
    config4 = None
    if not args.no_config4:
        # the learner (rank 0) runs on a side stream UNDER the rollout kernels of the same iteration (rollout.AsyncTrainer: 2 SMs
        # are left free for its 14 + 2 small dependent kernels; it samples only rows that earlier iterations completed; its policy is
        # used one iteration later).  With world > 1 the side stream also carries the broadcast of the flat parameter vector
        # (526 KB, NCCL) before the re-pack into the spare packed actor, and the statistics all-reduce stays inside the step.
        def run_c4(reserve):
            tr = tt.AsyncTrainer(eng, reserve_sms=reserve, sync=sync if world > 1 else None, is_learner=rank == 0)

            def c4_step():
                tr.step()
                if world > 1:
                    sync.push_stats(env.stats_tensor(clear=True))
            for _ in range(max(W, 5) + 60):
                c4_step()
            l0 = L.tt_launch_count()
            ms4, _ = timed(c4_step, K)
            n_l = int(L.tt_launch_count() - l0)
            # the rollout alone right behind it: same clock / thermal state, same reserved SMs (the headline region ran much earlier)
            torch.cuda.current_stream().wait_stream(tr.side)
            for _ in range(max(W, 5) + 20):
                headline_step()
            ms_plain, _ = timed(headline_step, K)
            tr.close()
            if world > 1:
                torch.cuda.current_stream().wait_stream(sync.stream)
                sync.flush_stats()
            return ms4, n_l, tr.updates, ms_plain
        ms4, n_l, n_upd, ms_plain = run_c4(2)
        # for reference: the same iteration with the update SERIAL on the rollout stream (no SMs reserved)
        agent.actor = actors[0]
        ln = agent.learner

        def c4_serial():
            eng.step(); ln.learn()
        for _ in range(max(W, 5)):
            c4_serial()
        ms4s, _ = timed(c4_serial, K)
        config4 = {"value": world * N * K / (ms4 * 1e-3), "unit": UNIT, "ms_per_step": ms4 / K, "gpu_launches": n_l,
                   "extra_ms_over_rollout": ms4 / K - ms_plain / K, "rollout_ms_per_step_same_state": ms_plain / K,
                   "extra_ms_over_headline": ms4 / K - ms / K, "reserved_sms": 2, "learner_updates_so_far": n_upd,
                   "serial_ms_per_step": ms4s / K,
                   "note": "configs[3]: rollout + fused replay store + one DDPG update per iteration (batch 64, tt_learn_step: hand-written "
                           "kernels on the device ring, hidden on a side stream under the rollout kernels, policy used one iteration later"
                           + (", learner on rank 0; stats all-reduce + broadcast of the flat actor vector over NCCL on side streams)" if world > 1 else ")")
                           + "; serial_ms_per_step = the update on the rollout stream instead; extra_ms_over_rollout is against rollout_ms_per_step_same_state "
                             "(the rollout alone timed right behind the config4 region, same reserved SMs)"}
        agent.actor = actors[0]

    # ---- BASELINE.json configs[4]: mini-sweep over envs per GPU (the full sweep: profiles/sweep.py) ----
    sweep = None
    if not args.no_sweep:
        sweep = []
        for n in (1 << 12, 1 << 18):
            e2 = tt.VecTruckTrailerEnv(n, seed=27, global_env_offset=rank * n, device=dev)
            a2 = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=max(4 * n, 1 << 16), num_envs=n, device=dev, seed=27,
                             global_env_offset=rank * n, precision=args.precision, actor_seed=0, allow_out_of_bar=args.precision == "bf16")
            g2 = tt.RolloutEngine(e2, a2, store=True)
            g2.reset()
            for _ in range(args.preroll):
                g2.step()
            iters = 200 if n <= (1 << 14) else 50
            l0 = L.tt_launch_count()
            ms_s, _ = timed(g2.step, iters)
            row = {"envs_per_gpu": n, "us_per_iteration": 1e3 * ms_s / iters, "value": world * n * iters / (ms_s * 1e-3),
                   "launches_per_iteration": (L.tt_launch_count() - l0) / iters}
            if n <= (1 << 14):                                       # launch-bound sizes: K iterations as ONE CUDA graph
                kk = g2.capture()
                reps = max(1, iters // kk)
                ms_g, _ = timed(g2.step_graph, reps)
                row["us_per_iteration_cuda_graph"] = 1e3 * ms_g / (reps * kk)
            sweep.append(row)
            del g2, a2, e2
        sweep.append({"envs_per_gpu": N, "us_per_iteration": 1e3 * ms / K, "value": value, "launches_per_iteration": launches / K})

    # ---- CPU baseline (rank 0, N = 1 only): the reference's Python loop on the host cores + the oracle's C port ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(args)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16": "bf16", "f16": "f16"}[args.precision],
                "data": "synthetic",
                "config": {"workload": WORKLOAD.replace("2^22 envs/GPU", f"{N} envs/GPU") + f"; actor {args.precision} (tcgen05), env step f64/f32 DP5",
                           "envs_per_gpu": N, "ring_capacity": args.ring, "preroll_iterations": args.preroll,
                           "launches_per_iteration": launches / K,
                           "l2": "working set per step (>1.2 GB/GPU) far exceeds the 126 MB L2; no flush needed",
                           "sharding": "global env id ranges, no data-path collective" + ("; the per-iteration statistics all-reduce (NCCL, side stream) "
                                       "is inside the timed step" if world > 1 else "")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
                "cpu_baseline": cpu, "rollout_stats": ttd.summarize(stats)}
        if config4 is not None:
            line["config4"] = config4
        if sweep is not None:
            line["sweep"] = sweep
        print(json.dumps(line), flush=True)
    if sampler:
        sampler.stop(0, 0)
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
