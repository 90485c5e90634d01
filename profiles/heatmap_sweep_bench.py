"""Row f2: the reference's heat-map data generation (DDPG/heatmap.py:39-193, "This may take several hours") as ONE batch.
Times `evaluate.run_sweep` on the reference's default grid (30 x 15 cells x 3 trials) and on a 16x finer grid, and puts the
reference's own per-step cost next to it (env.step + choose_action of the unmodified Python loop, measured by bench.py's cpu_baseline
on the same class of host: 0.61 ms per env step in one process).
    python profiles/heatmap_sweep_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import evaluate

REF_MS_PER_STEP = 0.61
for res, label in ((2.0, "reference grid (2 m cells)"), (0.5, "16 x finer grid (0.5 m cells)")):
    poses = evaluate.heatmap_poses(resolution=res)
    n = poses["start_x"].size
    agent = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=1, precision="f16", actor_seed=0)
    evaluate.run_sweep(agent, poses)                      # warm-up (kernel attributes, allocations)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = evaluate.run_sweep(agent, poses)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    steps = int(np.sum(out["trials"]["episode_steps"])) if "episode_steps" in out.get("trials", {}) else None
    msg = f"{label}: {n} episodes in one batch: {dt * 1e3:.0f} ms"
    if steps:
        msg += f" ({steps} env steps; the reference's sequential loop at {REF_MS_PER_STEP} ms per step: {steps * REF_MS_PER_STEP / 1e3:.0f} s)"
    print(msg)
