"""Batched evaluation sweep = the reference's heat-map data generation.

``DDPG/heatmap.py:39-193`` (``generate_heatmap_data``) walks a grid of trailer start positions one episode at a time
("This may take several hours"): per trial it draws a start heading and a trailer length (``env.L2 = uniform(5, 7)``,
heatmap.py:86-89), injects the start state with the truck L2 ahead of the trailer (heatmap.py:113-119), rolls the policy
out with ``evaluate=True`` until ``done`` and records the return, the success flag, the termination class and the
trailer end point.  Here every (cell, trial) pair is one environment of ONE batch: state and L2 injection through
``tt_env_set_state`` / ``tt_env_set_l2``, noise-free actor, finished envs freeze (reward 0), so the whole sweep is
``max_episode_steps`` launches of the rollout kernels.
"""
from __future__ import annotations

import numpy as np
import torch

# heatmap.py:20-31 defaults
GRID_RESOLUTION = 2.0
MAP_X_RANGE = (-30.0, 30.0)
MAP_Y_RANGE = (0.0, 30.0)
START_ORIENTATION_RANGE_DEG = (60.0, 120.0)
TRIALS_PER_CELL = 3
TERMINATION_CLASSES = ("success", "jackknife", "out_of_map", "goal_passed", "max_steps", "other_failure")


def classify(flags, success):
    """heatmap.py:158-172: success | jackknife | out_of_map | goal_passed | max_steps | other_failure, in that order of
    precedence, from the env's termination bits (TT_F_*) and info['success']."""
    flags = np.asarray(flags, np.uint8)
    out = np.full(flags.shape, 5, np.int64)
    out[(flags & 4) != 0] = 4          # max_steps_reached
    out[(flags & 16) != 0] = 3         # goal_passed
    out[(flags & 2) != 0] = 2          # out_of_map
    out[(flags & 1) != 0] = 1          # jackknife
    out[((flags & 8) != 0) | np.asarray(success, bool)] = 0
    return out


def heatmap_poses(x_range=MAP_X_RANGE, y_range=MAP_Y_RANGE, resolution=GRID_RESOLUTION, trials=TRIALS_PER_CELL,
                  yaw_range_deg=START_ORIENTATION_RANGE_DEG, l2_range=(5.0, 7.0), seed=0):
    """The sweep's trial list (heatmap.py:52-53,77-89): grid of start positions x ``trials`` draws of heading and L2."""
    rng = np.random.default_rng(seed)
    xs, ys = np.arange(x_range[0], x_range[1], resolution), np.arange(y_range[0], y_range[1], resolution)
    gx, gy = np.meshgrid(xs, ys)                                        # [ny, nx], row = y like reward_grid[iy, ix]
    sx = np.repeat(gx.reshape(-1), trials); sy = np.repeat(gy.reshape(-1), trials)
    yaw_deg = rng.uniform(yaw_range_deg[0], yaw_range_deg[1], sx.size)
    l2 = rng.uniform(l2_range[0], l2_range[1], sx.size)
    return dict(x_coords=xs, y_coords=ys, start_x=sx, start_y=sy, yaw_deg=yaw_deg, yaw_rad=np.deg2rad(yaw_deg), L2=l2, trials=trials)


def start_states(start_x, start_y, yaw_rad, l2):
    """heatmap.py:113-119: psi1 = psi2 = yaw, truck L2 ahead of the trailer, float32 state."""
    st = np.stack([yaw_rad, yaw_rad, start_x + l2 * np.cos(yaw_rad), start_y + l2 * np.sin(yaw_rad), start_x, start_y], 1)
    return st.astype(np.float32).astype(np.float64)


@torch.no_grad()
def run_sweep(agent, poses, goal=None, precision=None, record_trajectories=False, device=None, max_steps_cap=400):
    """Roll the policy out (``evaluate=True``: no noise) from every pose of ``poses`` (see heatmap_poses) in one batch.

    Returns the reference's outputs (heatmap.py:193): ``reward_grid``, ``success_grid`` [ny, nx] (mean over the trials of a
    cell), ``x_coords``, ``y_coords``, ``orientations`` (list of dicts), ``trajectory_endpoints`` (list of dicts with
    end_x/end_y/start_x/start_y/violation_type/score) and, when ``record_trajectories``, the trailer path of the first
    trial of each cell; plus flat per-trial arrays under ``trials``.
    """
    from .env import VecTruckTrailerEnv
    from .agent import VecAgent
    n = poses["start_x"].size
    dev = device if device is not None else agent.device
    env = VecTruckTrailerEnv(n, seed=0, emit_info=True, device=dev)
    env.set_l2(poses["L2"])
    st = start_states(poses["start_x"], poses["start_y"], poses["yaw_rad"], poses["L2"])
    start = np.stack([poses["start_x"], poses["start_y"], poses["yaw_rad"]], 1)
    gl = None if goal is None else np.tile(np.asarray(goal, np.float64), (n, 1))
    obs = env.set_state(st, start, gl)
    prec = precision or agent.precision
    score = torch.zeros(n, dtype=torch.float64, device=dev)
    alive = torch.ones(n, dtype=torch.bool, device=dev)
    flags = torch.zeros(n, dtype=torch.uint8, device=dev)
    success = torch.zeros(n, dtype=torch.bool, device=dev)
    ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
    first = np.arange(0, n, poses["trials"])
    traj = [[st[first, 4].copy()], [st[first, 5].copy()]] if record_trajectories else None
    tlen = np.ones(len(first), np.int64)
    steps = 0
    while bool(alive.any()) and steps < max_steps_cap:
        mu = agent.actor.forward(obs, precision=prec, allow_out_of_bar=getattr(agent, "allow_out_of_bar", False))                  # choose_action(obs, evaluate=True), heatmap.py:139
        obs, rew, done, info = env.step(VecAgent.scale_action(mu))     # heatmap.py:140-141
        score += torch.where(alive, rew.double(), torch.zeros_like(score))
        ep_len += alive
        newly = alive & done
        flags = torch.where(newly, info["termination_flags"], flags)
        success = torch.where(newly, info["success"], success)
        if record_trajectories:
            s = env.state[first].cpu().numpy()
            a = alive[first].cpu().numpy()
            traj[0].append(s[:, 4]); traj[1].append(s[:, 5]); tlen += a
        alive = alive & ~done
        steps += 1
    end = env.state.cpu().numpy()
    sc, fl, su = score.cpu().numpy(), flags.cpu().numpy(), success.cpu().numpy()
    cls = classify(fl, su)
    ny, nx, tr = len(poses["y_coords"]), len(poses["x_coords"]), poses["trials"]
    out = dict(reward_grid=sc.reshape(ny, nx, tr).mean(2), success_grid=su.reshape(ny, nx, tr).mean(2).astype(np.float64),
               x_coords=poses["x_coords"], y_coords=poses["y_coords"],
               orientations=[{"x": float(x), "y": float(y), "yaw_deg": float(d), "yaw_rad": float(r)}
                             for x, y, d, r in zip(poses["start_x"], poses["start_y"], poses["yaw_deg"], poses["yaw_rad"])],
               trajectory_endpoints=[{"end_x": float(end[i, 4]), "end_y": float(end[i, 5]), "start_x": float(poses["start_x"][i]),
                                      "start_y": float(poses["start_y"][i]), "violation_type": TERMINATION_CLASSES[cls[i]],
                                      "score": float(sc[i])} for i in range(n)],
               trials=dict(score=sc, success=su, flags=fl, termination_class=cls, end_state=end, steps=steps,
                           episode_steps=ep_len.cpu().numpy()))
    if record_trajectories:
        tx, ty = np.stack(traj[0], 1), np.stack(traj[1], 1)
        out["trajectories"] = [{"trailer_x": tx[c, :tlen[c]].tolist(), "trailer_y": ty[c, :tlen[c]].tolist(),
                                "start_x": float(poses["start_x"][first[c]]), "start_y": float(poses["start_y"][first[c]]),
                                "start_yaw_deg": float(poses["yaw_deg"][first[c]]), "success": bool(su[first[c]])} for c in range(len(first))]
    return out
