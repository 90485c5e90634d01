"""TEST INFRASTRUCTURE: golden vectors for the per-env trailer length.

Runs the UNTOUCHED reference env the way DDPG/heatmap.py:85-150 does -- ``env.L2 = uniform(5, 7)``, start pose on a
grid, float32 start state with the truck L2 ahead of the trailer, ``compute_observation(state, 0.0)`` -- and records
whole episodes under smooth / small steering.  Output: tests/golden/ref_l2.npz.  Needs /root/reference (build container
only); the fixture is committed.

    python oracle/make_golden_l2.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as rh           # noqa: E402
from oracle.make_golden import _actions, GOAL  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main(seed=77, n=32, T=260):
    rng = np.random.default_rng(seed)
    env = rh.make_env()
    eps = []
    for e in range(n):
        L2 = float(rng.uniform(5, 7))
        sx, sy = float(rng.uniform(-25, 25)), float(rng.uniform(2, 27))
        syaw = float(np.deg2rad(rng.uniform(60, 120)))
        env.L2 = L2                                             # heatmap.py:89
        st = np.array([syaw, syaw, sx + L2 * np.cos(syaw), sy + L2 * np.sin(syaw), sx, sy], dtype=np.float32)   # heatmap.py:116-119
        acts = _actions("smooth" if e % 2 else "small", T, rng)
        out = rh.rollout_reference(env, st.astype(np.float64), (sx, sy, syaw), GOAL, acts)
        out.update(L2=L2, state0=st.astype(np.float64), start=np.array([sx, sy, syaw]), actions=acts)
        eps.append(out)
    E, Tm = len(eps), max(len(e["done"]) for e in eps)
    pad = lambda key, shape, dt: np.stack([np.concatenate([e[key], np.zeros((Tm - len(e[key]),) + shape, dt)]) for e in eps])
    np.savez_compressed(
        os.path.join(OUT, "ref_l2.npz"),
        L2=np.array([e["L2"] for e in eps]), state0=np.stack([e["state0"] for e in eps]), start=np.stack([e["start"] for e in eps]),
        goal=np.tile(np.asarray(GOAL, np.float64), (E, 1)), actions=np.stack([e["actions"] for e in eps]),
        length=np.array([len(e["done"]) for e in eps]), obs0=np.stack([e["obs0"] for e in eps]),
        max_steps=np.array([e["max_steps"] for e in eps]),
        state=pad("state", (6,), np.float64), obs=pad("obs", (23,), np.float32), comps=pad("comps", (11,), np.float64),
        viol=pad("viol", (), np.uint8), flags=pad("flags", (6,), np.uint8), done=pad("done", (), np.uint8), success=pad("success", (), np.uint8))
    print("episodes", E, "lengths", [len(e["done"]) for e in eps])


if __name__ == "__main__":
    main()
