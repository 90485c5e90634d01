"""Parity of the fused env kernel (through the C ABI) with the float64 oracle and the reference's golden
vectors.  Bars (BASELINE.json north_star): state trajectories within 1e-4 relative (abs floor 1: the state
components cross 0), rewards within 1e-4 relative, termination flags exact except documented threshold
crossings within epsilon."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BITS = np.array([1, 2, 4, 8, 16, 32])


def _tt():
    import ddpg_trucktrailer_b200 as tt
    return tt


def _rel(a, b):
    return (np.abs(a - b) / np.maximum(np.abs(b), 1.0)).max()


def test_recorded_episode_10579(golden_dir):
    """The reference's own recorded episode: 193 steps, success, return 4792.9998."""
    tt = _tt()
    g = np.load(os.path.join(golden_dir, "episode_10579.npz"))
    env = tt.VecTruckTrailerEnv(1, emit_info=True)
    env.set_state(g["states"][0][None], g["start"][None], g["goal"][None])
    ret, states = 0.0, []
    for t, a in enumerate(g["actions"]):
        obs, rew, done, info = env.step(torch.tensor([a], device="cuda"))
        states.append(env.state[0].cpu().numpy())
        comps = np.array([float(rew[0])] + [float(info[k][0]) for k in tt.COMP_NAMES])
        assert np.abs(comps - g["comps"][t]).max() < 5e-4, t
        assert int(info["violation_type"][0]) == g["viol"][t]
        assert bool(done[0]) == (t == 192)
        ret += float(rew[0])
    assert bool(info["success"][0]) and abs(ret - 4792.9998) < 2e-2
    states = np.array(states)
    assert _rel(states, g["states"][1:]) < 1e-5
    assert np.abs(states[:, :2] - g["states"][1:, :2]).max() < 1e-7


def test_reference_rollouts_batched(golden_dir):
    """125 reference episodes (every termination type / violation code) as ONE batch, stepped in lock-step;
    finished envs freeze.  Also exercises ld_obs = 24."""
    tt = _tt()
    R = np.load(os.path.join(golden_dir, "ref_rollouts.npz"))
    E, T = R["actions"].shape
    env = tt.VecTruckTrailerEnv(E, emit_info=True, ld_obs=24)
    obs0 = env.set_state(R["state0"], R["start"], R["goal"])
    assert np.abs(obs0.cpu().numpy() - R["obs0"]).max() < 5e-7
    assert np.array_equal(env.get_state()["max_episode_steps"].cpu().numpy(), R["max_steps"])
    length = R["length"]
    acts = torch.from_numpy(R["actions"]).cuda()
    for t in range(T):
        obs, rew, done, info = env.step(acts[:, t])
        live = np.nonzero(length > t)[0]
        st = env.state.cpu().numpy()
        assert _rel(st[live], R["state"][live, t]) < 1e-5, t
        assert np.abs(obs.cpu().numpy()[live] - R["obs"][live, t]).max() < 5e-6, t
        comps = np.concatenate([rew.cpu().numpy()[:, None], torch.stack([info[k] for k in tt.COMP_NAMES], 1).cpu().numpy()], 1)
        assert _rel(comps[live], R["comps"][live, t]) < 1e-4, t
        assert np.array_equal(done.cpu().numpy()[live].astype(np.uint8), R["done"][live, t]), t
        assert np.array_equal(info["termination_flags"].cpu().numpy()[live], (R["flags"][live, t] * BITS).sum(1)), t
        assert np.array_equal(info["violation_type"].cpu().numpy()[live], R["viol"][live, t]), t
        assert np.array_equal(info["success"].cpu().numpy()[live].astype(np.uint8), R["success"][live, t]), t
        # finished earlier and not reset -> frozen: reward 0, done 1
        frozen = np.nonzero((length <= t) & (R["done"][np.arange(E), length - 1] == 1))[0]
        assert (rew.cpu().numpy()[frozen] == 0).all() and done.cpu().numpy()[frozen].all()


def _oracle_batch(orc, state0, start, actions):
    """Oracle trajectories for a batch: returns per-step arrays padded after done."""
    N, T = actions.shape
    out = dict(state=np.zeros((N, T, 6)), rew=np.zeros((N, T)), done=np.zeros((N, T), np.uint8), length=np.zeros(N, np.int32),
               flags=np.zeros((N, T), np.uint8), obs_last=np.zeros((N, 23), np.float32), margins=np.zeros((N, T, orc.NMARGINS)))
    cfg = orc.default_cfg(0)
    for i in range(N):
        e = orc.OracleEnv(cfg)
        e.set_state(state0[i], start[i])
        o = e.replay_m(actions[i])
        n = len(o["done"])
        out["state"][i, :n] = o["state"]; out["rew"][i, :n] = o["comps"][:, 0]; out["done"][i, :n] = o["done"]
        out["margins"][i, :n] = o["margins"]
        out["flags"][i, :n] = (o["flags"] * BITS).sum(1); out["length"][i] = n; out["obs_last"][i] = o["obs"][-1]
    return out


@pytest.mark.parametrize("N,T", [(4096, 120), (65536, 100)])
def test_against_oracle_random_batch(N, T):
    """BASELINE.json configs[1] shape (fixed action replay vs reference trajectories) at a size the oracle
    finishes in seconds: Philox start poses, half uniform / half smooth steering, K steps in ONE launch."""
    tt = _tt()
    from oracle import oracle as orc
    rng = np.random.default_rng(1234)
    env = tt.VecTruckTrailerEnv(N, seed=77, emit_info=False)
    env.reset()
    s = env.get_state()
    state0, start = s["state"].cpu().numpy(), s["start"].cpu().numpy()
    # reset parity: poses bit-exact with the shared Philox spec, float32-rounded states exact
    for i in (0, 1, N // 2, N - 1):
        assert tuple(start[i]) == orc.rng_pose(77, i, 0x80000000)
        e = orc.OracleEnv(); e.reset_pose(*start[i])
        assert np.array_equal(e.state[:2], state0[i][:2]) and np.abs(e.state[2:] - state0[i][2:]).max() <= 2.0 ** -26   # fixed-point positions
    acts = np.zeros((N, T), np.float32)
    acts[: N // 2] = rng.uniform(-np.pi / 4, np.pi / 4, (N // 2, T))
    walk = np.cumsum(rng.normal(0, 0.08, (N - N // 2, T)), 1)
    acts[N // 2:] = np.clip(walk, -np.pi / 4, np.pi / 4)
    obs, rew, done, info = env.step_k(torch.from_numpy(acts.T.copy()).cuda(), auto_reset=False, want_obs=True, want_info=True)
    rew, done = rew.cpu().numpy().T, done.cpu().numpy().T
    flags = info.flags.cpu().numpy().T
    ref = _oracle_batch(orc, state0, start, acts)
    L = ref["length"]
    # done-step agreement: termination flags are exact, except where the float64 oracle state sits within the documented
    # epsilon of the threshold whose flag differs (oracle.MARGIN_EPS) -- every mismatch is checked
    # against the oracle's margins at the step where the two paths part, none is accepted by count
    first_done = np.where(done.any(1), done.argmax(1) + 1, T + 1)
    ref_done = np.where(ref["done"].any(1), L, T + 1)
    mism = np.nonzero(first_done != ref_done)[0]
    for i in mism:
        t = min(first_done[i], ref_done[i]) - 1
        assert orc.explained_by_margin(flags[i, t], ref["flags"][i, t], ref["margins"][i, t]), \
            f"env {i} step {t}: flags {flags[i, t]:#x} vs oracle {ref['flags'][i, t]:#x}, margins {ref['margins'][i, t]} -- not a within-epsilon crossing"
    assert len(mism) <= max(1, N // 2000), f"{len(mism)} epsilon crossings is more than a float32-level effect"
    ok = np.setdiff1d(np.arange(N), mism)
    mask = np.arange(T)[None, :] < L[:, None]
    mask[mism] = False
    assert _rel(rew[mask], ref["rew"][mask]) < 1e-4
    assert np.array_equal(flags[mask], ref["flags"][mask])
    final = env.state.cpu().numpy()
    ref_final = ref["state"][np.arange(N), L - 1]
    assert _rel(final[ok], ref_final[ok]) < 1e-5
    # terminal observation rows written by the K-step kernel
    last_obs = obs.cpu().numpy()[np.minimum(first_done, T) - 1, np.arange(N)]
    assert np.abs(last_obs[ok] - ref["obs_last"][ok]).max() < 1e-5


def test_step_k_equals_single_steps_and_global_ids():
    """(1) K steps in one launch == K single launches (auto-reset on, bit-exact).  (2) multi-GPU invariance:
    an env's trajectory depends on (seed, GLOBAL id) only, so a shard with offset equals the slice."""
    tt = _tt()
    N, K = 1000, 40
    acts = torch.from_numpy(np.random.default_rng(0).uniform(-0.8, 0.8, (K, N)).astype(np.float32)).cuda()
    a = tt.VecTruckTrailerEnv(N, seed=5); a.reset()
    b = tt.VecTruckTrailerEnv(N, seed=5); b.reset()
    _, rk, dk, _ = a.step_k(acts, auto_reset=True)
    rs, ds = [], []
    for t in range(K):
        _, r, d, _ = b.step(acts[t]); rs.append(r.clone()); ds.append(d.clone())
        b.reset(options={"mask": d})
    assert torch.equal(rk, torch.stack(rs)) and torch.equal(dk.bool(), torch.stack(ds))
    assert torch.equal(a.state, b.state) and dk.sum() > 50
    c = tt.VecTruckTrailerEnv(300, seed=5, global_env_offset=600); c.reset()
    _, rc, dc, _ = c.step_k(acts[:, 600:900].contiguous(), auto_reset=True)
    assert torch.equal(rc, rk[:, 600:900]) and torch.equal(c.state, a.state[600:900])


def test_auto_reset_distribution_and_stats():
    """Philox reset poses follow simv2.py:331-333; device statistics count what happened."""
    tt = _tt()
    N, K = 1 << 16, 64
    env = tt.VecTruckTrailerEnv(N, seed=27); env.reset()
    s = env.get_state()
    st = s["start"].cpu().numpy()
    lo, hi = np.array([-27, 0, np.deg2rad(45)]), np.array([27, 27, np.deg2rad(120)])
    u = (st - lo) / (hi - lo)
    assert (u >= 0).all() and (u < 1).all()
    assert np.abs(u.mean(0) - 0.5).max() < 0.01 and np.abs(u.var(0) - 1 / 12).max() < 0.003
    acts = torch.empty(K, N, device="cuda").uniform_(-0.785, 0.785)
    _, rew, done, _ = env.step_k(acts, auto_reset=True)
    stats = env.read_stats()
    assert stats["steps"] == N * K and stats["episodes"] == int(done.sum())
    assert abs(stats["reward_sum"] - float(rew.double().sum())) < 1e-3 * abs(float(rew.double().sum())) + 1
    assert sum(stats["term_" + f] for f in tt.FLAG_NAMES) >= stats["episodes"] > N // 4
    st2 = env.get_state()["start"].cpu().numpy()
    changed = (st2 != st).any(1)
    assert changed.sum() >= 0.9 * min(N, stats["episodes"]) * 0.5


def test_n1_reference_contract(golden_dir):
    """Truck_trailer_Env_2 (N=1 wrapper) satisfies the caller contract of DDPG/trainv2.py:488-531."""
    tt = _tt()
    env = tt.Truck_trailer_Env_2(seed=27)
    obs, info = env.reset(seed=27 + 3)
    assert isinstance(obs, np.ndarray) and obs.shape == (23,) and obs.dtype == np.float32 and info == {}
    assert env.action_space.high.dtype == np.float32 and env.action_space.shape[0] == 1 and env.observation_space.shape == (23,)
    assert env.reward_range[0] == -float("inf")
    s0 = env.state.copy()
    assert s0.shape == (6,) and s0[0] == s0[1] and env.episode_steps == 0 and env.max_episode_steps == env.compute_max_steps()
    score, done, n = 0.0, False, 0
    while not done:
        scaled = np.clip(np.array([0.3], np.float32), -1, 1) * env.action_space.high
        obs_, reward, done, info = env.step(scaled)
        assert isinstance(reward, float) and isinstance(done, bool) and isinstance(info, dict)
        for k in ("final_success_bonus", "success", "violation_type", "total_reward", "progress_reward"):
            assert k in info
        score += reward; n += 1
    assert info["violation_type"] in tt.VIOLATION_NAMES and n == env.episode_steps
    # state injection like DDPG/test.py:96-115
    g = np.load(os.path.join(golden_dir, "episode_10579.npz"))
    env.startx, env.starty, env.startyaw = g["start"]
    env.state = g["states"][0]
    o, r, d, i = env.step(g["actions"][:1])
    assert abs(r - g["comps"][0, 0]) < 1e-3 and np.abs(env.state - g["states"][1]).max() < 1e-5
    # all 14 keys of the reference's info dict (reward_functionv1.py:488-504), incl. the backward-movement diagnostics
    want = {"total_reward", "distance_reward", "progress_reward", "heading_reward", "orientation_reward", "staged_success", "safety_penalty",
            "exploration_bonus", "final_success_bonus", "violation_type", "backward_penalty", "smoothness_penalty", "backward_movement_info", "success"}
    assert want <= set(i)
    for t in range(1, 60):
        o, r, d, i = env.step(g["actions"][t:t + 1])
        assert abs(i["backward_movement_info"]["cumulative_backward"] - g["cumulative_backward"][t]) < 1e-4, t
        assert i["backward_movement_info"]["movement_budget"] == pytest.approx(5.0 * min(1.0, (t + 1) / 50))


def test_full_size_properties():
    """BASELINE.json's full size (2^22 envs on one GPU): size-independent properties instead of an oracle replay.
    (1) K steps in one launch == K single launches, bit for bit, incl. auto-reset; (2) the device statistics equal
    the per-step outputs; (3) every env that reports done was reset (episode_steps == 0) and only those;
    (4) a shard with a global offset reproduces the corresponding slice."""
    tt = _tt()
    N, K = 1 << 22, 6
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    acts = torch.empty(K, N, device="cuda").uniform_(-0.785, 0.785, generator=g)
    a = tt.VecTruckTrailerEnv(N, seed=9); a.reset()
    b = tt.VecTruckTrailerEnv(N, seed=9); b.reset()
    # warm both populations so that episodes end inside the window
    warm = torch.empty(60, N, device="cuda").uniform_(-0.785, 0.785, generator=g)
    a.step_k(warm, auto_reset=True, want_reward=False, want_done=False); b.step_k(warm, auto_reset=True, want_reward=False, want_done=False)
    a.read_stats(); b.read_stats()
    _, rk, dk, _ = a.step_k(acts, auto_reset=True)
    rsum, dsum = 0.0, 0
    for t in range(K):
        _, r, d, _ = b.step(acts[t])
        assert torch.equal(r, rk[t]) and torch.equal(d, dk[t].bool())
        rsum += float(r.double().sum()); dsum += int(d.sum())
        steps_before = b.get_state()["episode_steps"]
        b.reset(options={"mask": d})
        steps_after = b.get_state()["episode_steps"]
        assert (steps_after[d] == 0).all() and torch.equal(steps_after[~d], steps_before[~d])
    assert torch.equal(a.state, b.state)
    sa, sb = a.read_stats(), b.read_stats()
    for k in sa:      # counts are exact; float sums depend on the (atomic) summation order
        if k in ("return_sum", "return_sq_sum", "reward_sum"):
            assert abs(sa[k] - sb[k]) <= 1e-5 * abs(sb[k]) + 1.0, k
        else:
            assert sa[k] == sb[k], k
    assert sa["steps"] == N * K and sa["episodes"] == dsum and dsum > N // 100
    assert abs(sa["reward_sum"] - rsum) <= 1e-6 * abs(rsum) + 1.0
    assert sum(sa["term_" + f] for f in tt.FLAG_NAMES) >= sa["episodes"]
    c = tt.VecTruckTrailerEnv(4096, seed=9, global_env_offset=N - 4096); c.reset()
    c.step_k(warm[:, N - 4096:].contiguous(), auto_reset=True, want_reward=False, want_done=False)
    _, rc, _, _ = c.step_k(acts[:, N - 4096:].contiguous(), auto_reset=True)
    assert torch.equal(rc, rk[:, N - 4096:])


def test_sparse_masked_reset_equals_per_env_reset():
    """The persistent mask-scanning reset kernel (N >= 2^16 with a 16 B aligned mask: warp-level compaction of the finished
    envs, one env per lane) resets exactly the masked envs and produces the same episodes / observations as the
    one-thread-per-env kernel (unaligned mask -> fallback), for sparse and for dense masks."""
    tt = _tt()
    N = (1 << 16) + 5                                   # ragged last group of 16
    for density in (0.02, 0.7):
        envs = [tt.VecTruckTrailerEnv(N, seed=31) for _ in range(2)]
        for e in envs:
            e.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        mask = (torch.rand(N, device="cuda", generator=g) < density).to(torch.uint8)
        mask[-1] = 1; mask[0] = 1
        raw = torch.zeros(N + 16, dtype=torch.uint8, device="cuda")
        unaligned = raw[1:N + 1]; unaligned.copy_(mask)
        assert unaligned.data_ptr() % 16 != 0 and mask.data_ptr() % 16 == 0
        before = envs[0].get_state()["state"].clone()
        a = torch.zeros(N, device="cuda")
        for e in envs:
            e.step(a)
        o0, _ = envs[0].reset(options={"mask": mask})
        o1, _ = envs[1].reset(options={"mask": unaligned})
        s0, s1 = envs[0].get_state(), envs[1].get_state()
        for k in ("state", "start", "goal", "episode_steps", "max_episode_steps"):
            assert torch.equal(s0[k], s1[k]), k
        assert torch.equal(o0, o1)
        m = mask.bool()
        assert (s0["episode_steps"][m] == 0).all() and (s0["episode_steps"][~m] == 1).all()
        assert not torch.equal(s0["state"][m], before[m])
