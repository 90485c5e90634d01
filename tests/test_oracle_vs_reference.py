"""Live differential test: C oracle vs the UNTOUCHED reference imported from /root/reference.
Runs only where the reference is mounted (the build container); skipped on the GPU box."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")


def test_rk45_matches_scipy_call_site():
    """solve_ivp(..., method='RK45') at simv2.py:510-515 vs the C restatement, incl. step counts."""
    env = rh.make_env()
    import scipy.integrate as spi
    rng = np.random.default_rng(0)
    nsteps = []
    for _ in range(300):
        y0 = np.array([rng.uniform(0, 3), rng.uniform(0, 3), *rng.uniform(-30, 30, 4)])
        y0[0] = y0[1] + rng.uniform(-1.4, 1.4)
        delta = float(rng.uniform(-np.pi / 4, np.pi / 4))
        env.steering_angle = delta
        sol = spi.solve_ivp(lambda t, y: env.kinematic_model(t, y, delta), [0, env.dt], y0, method="RK45")
        y, ns, nf = orc.rk45(y0, delta)
        assert ns == len(sol.t) - 1 and nf == sol.nfev
        assert np.abs(y - sol.y[:, -1]).max() < 1e-13
        nsteps.append(ns)
    assert max(nsteps) <= 3


def test_seeded_episodes_random_steering():
    """BASELINE.json configs[0]: 1 env, reset(seed), random steering, 1000 steps with driver-side resets."""
    env = rh.make_env()
    rng = np.random.default_rng(42)
    steps, ep = 0, 0
    while steps < 1000:
        obs, _ = env.reset(seed=500 + ep)
        e = orc.OracleEnv()
        o = e.reset_pose(env.startx, env.starty, env.startyaw)
        assert np.array_equal(e.state.astype(np.float32), env.state) and np.abs(o - obs).max() < 2.5e-7
        done = False
        while not done:
            a = np.float32(rng.uniform(-np.pi / 4, np.pi / 4))
            obs, rew, done, info = env.step(np.array([a], np.float32))
            o, comps, d, viol, flags, succ = e.step(a)
            assert d == bool(done)
            assert np.abs(e.state - env.state).max() < 1e-11
            assert np.abs(o - obs).max() < 2e-7
            assert abs(comps[0] - rew) < 5e-6     # float32 arctan2: numpy SIMD vs glibc, 1 ulp
            assert orc.VIOLATION_NAMES[viol] == info["violation_type"]
            steps += 1
        ep += 1
    assert ep > 5
