"""TEST INFRASTRUCTURE ONLY -- generates the committed golden fixtures under tests/golden/ by running the
UNTOUCHED reference (imported from /root/reference through oracle/ref_harness.py, build container only).

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

Fixtures (all small; consumed by tests/ on the CPU and on the GPU box, where /root/reference is absent):
  episode_10579.npz   the reference's own recorded episode DDPG/episode_replays/episode_10579_reward_4792.pkl
                      (193 steps: states, scaled actions, the reward components, success) re-packed as arrays.
  ref_rollouts.npz    reference trajectories for several action families chosen so that every termination
                      type (jackknife, out_of_map, max_steps, goal_reached, goal_passed, excessive_backward)
                      and every violation code occurs; ragged episodes padded to Tmax with `length`.
  ref_resets.npz      reference reset(seed) outputs: pose, float32 state, observation, max_episode_steps.
  ref_actor.npz       reference ActorNetwork (torch.manual_seed(0) init) state_dict, a "trained-like"
                      re-scaled copy, input observations and the reference torch forward outputs.
  ref_misc.npz        OUActionNoise trace for given normals; ReplayBuffer.store_transition wrap-around case.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GOAL = (0, -30, np.deg2rad(90))


def pack_golden_pkl():
    d = pickle.load(open(rh.golden_pkl_path(), "rb"))
    ed = d["env_data"]
    info = d["info"]
    np.savez_compressed(
        os.path.join(OUT, "episode_10579.npz"),
        states=np.array([np.asarray(s, np.float64) for s in d["states"]]),
        actions=np.array([a[0] for a in d["actions"]], np.float32),
        start=np.array([ed["startx"], ed["starty"], ed["startyaw"]], np.float64),
        goal=np.array([ed["goalx"], ed["goaly"], ed["goalyaw"]], np.float64),
        comps=np.array([[float(i[k]) for k in rh.REWARD_KEYS] for i in info], np.float64),
        viol=np.array([rh.VIOLATION_CODES[i["violation_type"]] for i in info], np.uint8),
        success=np.array([bool(i["success"]) for i in info], np.uint8),
        cumulative_backward=np.array([i["backward_movement_info"]["cumulative_backward"] for i in info]),
        episode_num=d["episode_num"])


def _actions(kind, T, rng):
    hi = np.float32(np.pi / 4)
    if kind == "uniform":
        a = rng.uniform(-np.pi / 4, np.pi / 4, T)
    elif kind == "smooth":
        a = np.clip(np.cumsum(rng.normal(0, 0.3 * np.pi / 4, T)), -np.pi / 4, np.pi / 4)
        a[0] = rng.uniform(-0.3, 0.3)
        for t in range(1, T):
            a[t] = np.clip(a[t - 1] + rng.normal(0, 0.3 * np.pi / 4), -np.pi / 4, np.pi / 4)
    elif kind == "small":
        a = rng.normal(0, 0.02, T)
    elif kind == "saturated":      # exceeds the float64 clip at simv2.py:504 (float32 pi/4 > float64 pi/4)
        a = np.where(rng.random(T) < 0.5, 1.5, -1.5) * np.where(rng.random(T) < 0.3, 1.0, 0.01)
    elif kind == "zero":
        a = np.zeros(T)
    else:
        raise ValueError(kind)
    return np.clip(a.astype(np.float32), -1.6, 1.6) if kind == "saturated" else np.minimum(np.maximum(a.astype(np.float32), -hi), hi)


def _pose_state(env, sx, sy, syaw, psi1=None, f32=True):
    x1 = sx + env.L2 * np.cos(syaw); y1 = sy + env.L2 * np.sin(syaw)
    st = np.array([syaw if psi1 is None else psi1, syaw, x1, y1, sx, sy], np.float64)
    if psi1 is not None:   # truck not aligned with the trailer: hitch point fixed, truck ahead of it along psi1
        st[2] = sx + env.L2 * np.cos(syaw); st[3] = sy + env.L2 * np.sin(syaw)
    return st.astype(np.float32).astype(np.float64) if f32 else st


def gen_rollouts(seed=1234):
    rng = np.random.default_rng(seed)
    env = rh.make_env()
    eps = []

    def add(state0, start, actions, tag):
        out = rh.rollout_reference(env, state0, start, GOAL, actions)
        out.update(state0=np.asarray(state0, np.float64), start=np.asarray(start, np.float64), actions=np.asarray(actions, np.float32), tag=tag)
        eps.append(out)

    # (A) reference start distribution (simv2.py:331-333), random / smooth / small steering
    for kind, n in (("uniform", 16), ("smooth", 24), ("small", 16), ("saturated", 6)):
        for _ in range(n):
            sx, sy, syaw = rng.uniform(-27, 27), rng.uniform(0, 27), rng.uniform(np.deg2rad(45), np.deg2rad(120))
            add(_pose_state(env, sx, sy, syaw), (sx, sy, syaw), _actions(kind, 300, rng), "A_" + kind)
    # (B) near the goal, roughly aligned: success / goal_passed / staged bonuses
    for _ in range(30):
        d = rng.uniform(0.3, 9.0)
        syaw = np.deg2rad(90) + rng.normal(0, 0.12)
        sx, sy = rng.normal(0, 0.25), -30 + d
        add(_pose_state(env, sx, sy, syaw), (sx, sy, syaw), _actions("small", 60, rng), "B_near")
    # (C) the recorded successful action sequence from perturbed initial states (long episodes)
    d = pickle.load(open(rh.golden_pkl_path(), "rb"))
    ed = d["env_data"]
    acts = np.array([a[0] for a in d["actions"]], np.float32)
    for k in range(6):
        st0 = np.asarray(d["states"][0], np.float64).copy()
        st0[:2] += rng.normal(0, 1e-5 * (k + 1), 2)
        add(st0, (ed["startx"], ed["starty"], ed["startyaw"]), acts, "C_golden_perturbed")
    # (D) out of the map: start near the left/right edge heading outwards, straight
    for _ in range(6):
        side = rng.choice([-1, 1])
        sx, sy = side * rng.uniform(33, 38), rng.uniform(5, 25)
        syaw = np.deg2rad(90) + side * np.deg2rad(rng.uniform(40, 60)) * -1
        add(_pose_state(env, sx, sy, syaw, f32=False), (sx, sy, syaw), _actions("small", 120, rng) * 0.2, "D_boundary")
    # (E) max_steps: injected start close to the goal (short budget) while the vehicle is far away
    for _ in range(4):
        sx, sy, syaw = rng.normal(0, 1), rng.uniform(20, 27), np.deg2rad(90) + rng.normal(0, 0.02)
        add(_pose_state(env, sx, sy, syaw), (rng.normal(0, 0.5), -29.0, syaw), _actions("zero", 140, rng), "E_max_steps")
    # (F) excessive backward: drive away from the goal (trailer pointing down, so reversing moves up)
    for _ in range(6):
        sx, sy, syaw = rng.uniform(-5, 5), rng.uniform(-10, 5), np.deg2rad(-90) + rng.normal(0, 0.1)
        add(_pose_state(env, sx, sy, syaw, f32=False), (sx, sy, syaw), _actions("small", 80, rng), "F_excessive_backward")
    # (G) non-zero initial hitch angle and large headings
    for _ in range(8):
        sx, sy, syaw = rng.uniform(-20, 20), rng.uniform(5, 25), rng.uniform(np.deg2rad(45), np.deg2rad(120))
        add(_pose_state(env, sx, sy, syaw, psi1=syaw + rng.uniform(-0.6, 0.6), f32=False), (sx, sy, syaw),
            _actions("smooth", 200, rng), "G_hitch")

    # (H) injected state beyond the +/-42 m band: major_boundary violation code (reward_functionv1.py:394-400)
    for _ in range(3):
        sx, sy, syaw = rng.choice([-1, 1]) * rng.uniform(42.3, 43.0), rng.uniform(5, 20), np.deg2rad(90)
        add(_pose_state(env, sx, sy, syaw, f32=False), (sx, sy, syaw), _actions("zero", 3, rng), "H_major_boundary")

    E = len(eps)
    Tmax = max(len(e["done"]) for e in eps)
    pad = lambda key, shape, dt: np.zeros((E, Tmax) + shape, dt)
    out = dict(length=np.array([len(e["done"]) for e in eps], np.int32),
               state0=np.array([e["state0"] for e in eps]), start=np.array([e["start"] for e in eps]),
               goal=np.tile(np.asarray(GOAL, np.float64), (E, 1)), obs0=np.array([e["obs0"] for e in eps]),
               max_steps=np.array([e["max_steps"] for e in eps], np.int32),
               tag=np.array([e["tag"] for e in eps]),
               actions=pad("actions", (), np.float32), state=pad("state", (6,), np.float64),
               obs=pad("obs", (23,), np.float32), comps=pad("comps", (11,), np.float64),
               viol=pad("viol", (), np.uint8), flags=pad("flags", (6,), np.uint8), done=pad("done", (), np.uint8),
               success=pad("success", (), np.uint8))
    for i, e in enumerate(eps):
        n = len(e["done"])
        out["actions"][i, :min(Tmax, len(e["actions"]))] = e["actions"][:Tmax]
        for k in ("state", "obs", "comps", "viol", "flags", "done", "success"):
            out[k][i, :n] = e[k]
    np.savez_compressed(os.path.join(OUT, "ref_rollouts.npz"), **out)
    fl = np.array([e["flags"][-1] for e in eps])
    print("episodes", E, "steps", int(out["length"].sum()), "Tmax", Tmax)
    print("terminal flags (jk, oom, max, reached, passed, exb):", fl.sum(0), " undone:", int(sum(1 - e["done"][-1] for e in eps)))
    print("violation codes seen:", sorted(set(np.concatenate([e["viol"] for e in eps]).tolist())))


def gen_resets():
    env = rh.make_env()
    seeds = np.arange(1000, 1064)
    rows = []
    for s in seeds:
        obs, _ = env.reset(seed=int(s))
        rows.append((env.startx, env.starty, env.startyaw, env.state.copy(), obs.copy(), env.max_episode_steps))
    np.savez_compressed(os.path.join(OUT, "ref_resets.npz"), seeds=seeds,
                        pose=np.array([[r[0], r[1], r[2]] for r in rows], np.float64),
                        state=np.array([r[3] for r in rows], np.float32), obs=np.array([r[4] for r in rows], np.float32),
                        max_steps=np.array([r[5] for r in rows], np.int32))


def gen_actor(seed=0):
    import torch
    actor = rh.make_actor(seed)
    sd = {k: v.detach().numpy().copy() for k, v in actor.state_dict().items()}
    rng = np.random.default_rng(7)
    # observations: rows of real reference observations + random rows in [-1, 1]
    R = np.load(os.path.join(OUT, "ref_rollouts.npz"))
    real = R["obs"][R["done"].astype(bool) | (np.arange(R["obs"].shape[1])[None, :] < R["length"][:, None])][:384]
    obs = np.concatenate([real, rng.uniform(-1, 1, (128, 23)).astype(np.float32)]).astype(np.float32)
    with torch.no_grad():
        out0 = actor.forward(torch.from_numpy(obs)).numpy()[:, 0]
        # "trained-like": larger last layer and non-trivial LayerNorm affine so tanh/LN are exercised
        sd2 = {k: v.copy() for k, v in sd.items()}
        sd2["mu.weight"] *= 60.0; sd2["mu.bias"] += 0.05
        sd2["bn1.weight"] = (1 + 0.3 * rng.standard_normal(400)).astype(np.float32)
        sd2["bn1.bias"] = (0.2 * rng.standard_normal(400)).astype(np.float32)
        sd2["bn2.weight"] = (1 + 0.3 * rng.standard_normal(300)).astype(np.float32)
        sd2["bn2.bias"] = (0.2 * rng.standard_normal(300)).astype(np.float32)
        sd2["fc2.weight"] *= 2.0
        actor.load_state_dict({k: torch.from_numpy(v) for k, v in sd2.items()})
        out1 = actor.forward(torch.from_numpy(obs)).numpy()[:, 0]
    save = {("w0/" + k): v for k, v in sd.items()}
    # the trained-like set is re-derived from w0 in the tests; only its deltas are stored
    for k in ("bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"):
        save["w1/" + k] = sd2[k]
    np.savez_compressed(os.path.join(OUT, "ref_actor.npz"), obs=obs, out0=out0, out1=out1, **save)
    print("actor fixture:", obs.shape, "out0 range", out0.min(), out0.max(), "out1 range", out1.min(), out1.max())


def gen_misc():
    _, noise_mod, rb_mod = rh.load_ddpg_modules()
    # OU noise (noise.py:12-17) driven by known normals: patch np.random.normal through a seeded stream
    np.random.seed(123)
    normals = np.random.normal(size=64)
    np.random.seed(123)
    ou = noise_mod.OUActionNoise(mu=np.zeros(1))
    trace = np.array([ou()[0] for _ in range(64)])
    # ReplayBuffer.store_transition (replay_buffer.py:13-21), capacity 10, 27 transitions -> wraps twice
    rb = rb_mod.ReplayBuffer(10, (23,), 1)
    rng = np.random.default_rng(5)
    s = rng.uniform(-1, 1, (27, 23)).astype(np.float32); s2 = rng.uniform(-1, 1, (27, 23)).astype(np.float32)
    a = rng.uniform(-1.2, 1.2, (27, 1)).astype(np.float32); r = rng.normal(0, 50, 27); d = rng.random(27) < 0.2
    for i in range(27):
        rb.store_transition(s[i], a[i], r[i], s2[i], d[i])
    np.savez_compressed(os.path.join(OUT, "ref_misc.npz"), ou_normals=normals, ou_trace=trace,
                        rb_s=s, rb_a=a, rb_r=r, rb_s2=s2, rb_d=d, rb_state=rb.state_memory, rb_new_state=rb.new_state_memory,
                        rb_action=rb.action_memory, rb_reward=rb.reward_memory, rb_terminal=rb.terminal_memory,
                        rb_cntr=rb.mem_cntr)


def main():
    os.makedirs(OUT, exist_ok=True)
    pack_golden_pkl()
    gen_rollouts()
    gen_resets()
    gen_actor()
    gen_misc()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
