"""GPU parity of the "next" rows of SURVEY.md section 8: f2 (per-env trailer length + batched evaluation sweep = the
reference's heat-map generation), f3 (episode / transition recording) and f4 (checkpoint interop)."""
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (np.abs(a - b) / np.maximum(np.abs(b), 1.0)).max()


def test_per_env_l2_matches_reference(golden_dir):
    """32 reference episodes run the heatmap.py way (env.L2 drawn per trial, state injected, heatmap.py:85-150) as ONE
    batch with per-env trailer lengths: north_star bars (state 1e-4 rel, reward 1e-4 rel, flags exact)."""
    import ddpg_trucktrailer_b200 as tt
    R = np.load(os.path.join(golden_dir, "ref_l2.npz"))
    E, T = R["actions"].shape
    env = tt.VecTruckTrailerEnv(E, emit_info=True)
    env.set_l2(R["L2"])
    assert np.array_equal(env.get_l2().cpu().numpy(), R["L2"]) and torch.is_tensor(env.L2)
    obs0 = env.set_state(R["state0"], R["start"], R["goal"])
    assert np.abs(obs0.cpu().numpy() - R["obs0"]).max() < 5e-7
    assert np.array_equal(env.get_state()["max_episode_steps"].cpu().numpy(), R["max_steps"])
    acts = torch.from_numpy(R["actions"]).cuda()
    length = R["length"]
    for t in range(int(length.max())):
        obs, rew, done, info = env.step(acts[:, t])
        live = np.nonzero(length > t)[0]
        st = env.state.cpu().numpy()
        assert _rel(st[live], R["state"][live, t]) < 1e-5, t
        assert np.abs(obs.cpu().numpy()[live] - R["obs"][live, t]).max() < 5e-6
        r, rr = rew.cpu().numpy()[live], R["comps"][live, t, 0]
        assert (np.abs(r - rr) / np.maximum(np.abs(rr), 1.0)).max() < 1e-4, t
        assert np.array_equal(done.cpu().numpy()[live], R["done"][live, t].astype(bool)), t
        assert np.array_equal(info["violation_type"].cpu().numpy()[live], R["viol"][live, t])
        fl = info["termination_flags"].cpu().numpy()[live]
        assert np.array_equal((fl[:, None] >> np.arange(6)) & 1, R["flags"][live, t])
    # the default trailer length gives different dynamics (the fixture is sensitive to L2)
    env2 = tt.VecTruckTrailerEnv(E)
    env2.set_state(R["state0"], R["start"], R["goal"])
    env2.step(acts[:, 0]); env2.step(acts[:, 1])
    assert np.abs(env2.state.cpu().numpy()[:, 1] - R["state"][:, 1, 1]).max() > 1e-4


def test_per_env_l2_reset_and_step_vs_oracle():
    """reset() places the truck L2 ahead of the trailer (simv2.py:483-484) and the dynamics use v/L2 (simv2.py:291):
    per-env lengths against the oracle configured with the same L2, including a subset injection by index."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    N = 64
    rng = np.random.default_rng(5)
    l2 = np.full(N, 7.0); idx = np.arange(0, N, 2); l2[idx] = rng.uniform(4.5, 8.0, idx.size)
    env = tt.VecTruckTrailerEnv(N, seed=11)
    env.set_l2(l2[idx], idx=idx)
    env.reset()
    s = env.get_state()
    st, sp = s["state"].cpu().numpy(), s["start"].cpu().numpy()
    want_x1 = (sp[:, 0] + l2 * np.cos(sp[:, 2])).astype(np.float32)
    assert np.abs(st[:, 2] - want_x1).max() < 1e-6 and np.abs(st[:, 4] - sp[:, 0].astype(np.float32)).max() < 1e-6
    acts = rng.uniform(-0.5, 0.5, (20, N)).astype(np.float32)
    for t in range(20):
        env.step(torch.from_numpy(acts[t]).cuda())
    got = env.state.cpu().numpy()
    for i in (0, 1, 2, 31, 62):
        cfg = orc.default_cfg(); cfg.L2 = float(l2[i])
        o = orc.OracleEnv(cfg)
        o.set_state(st[i], sp[i])
        for t in range(20):
            _, _, done, *_ = o.step(acts[t, i])
            if done:
                break
        if not done:
            assert _rel(got[i], o.state) < 1e-5, i


def test_n1_wrapper_l2_and_compute_observation(golden_dir):
    """heatmap.py:86-122 with the N=1 drop-in: env.L2 assignment, env.state assignment, compute_observation."""
    import ddpg_trucktrailer_b200 as tt
    R = np.load(os.path.join(golden_dir, "ref_l2.npz"))
    env = tt.Truck_trailer_Env_2()
    e = 3
    env.startx, env.starty, env.startyaw = R["start"][e]
    env.L2 = R["L2"][e]
    env.max_episode_steps = env.compute_max_steps()
    assert env.max_episode_steps == R["max_steps"][e]
    env.state = R["state0"][e].astype(np.float32)
    obs = env.compute_observation(env.state, steering_angle=0.0)
    assert obs.shape == (23,) and obs.dtype == np.float32 and np.abs(obs - R["obs0"][e]).max() < 5e-7
    obs_s = env.compute_observation(env.state, steering_angle=0.3)
    assert abs(obs_s[10] - np.sin(0.3)) < 1e-7 and abs(obs_s[11] - np.cos(0.3)) < 1e-7 and np.array_equal(np.delete(obs_s, [10, 11]), np.delete(obs, [10, 11]))
    score = 0.0
    for t in range(int(R["length"][e])):
        o, r, done, info = env.step(np.array([R["actions"][e, t]], np.float32))
        score += r
    assert done and abs(score - R["comps"][e, :R["length"][e], 0].sum()) < 1e-3 * max(1.0, abs(score))
    assert info["violation_type"] == tt.VIOLATION_NAMES[R["viol"][e, R["length"][e] - 1]]


def test_evaluation_sweep_vs_sequential_oracle(golden_dir):
    """generate_heatmap_data (heatmap.py:39-193) as one batch == the sequential loop over an oracle env + oracle actor."""
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import evaluate as ev
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    sd = {k[3:]: g[k].copy() for k in g.files if k.startswith("w0/")}
    sd["mu.weight"] = sd["mu.weight"] * np.float32(40.0)           # a policy that actually steers
    agent = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=1, max_size=64, precision="fp32")
    agent.load_actor_state_dict(sd)
    poses = ev.heatmap_poses(x_range=(-20, 20), y_range=(4, 24), resolution=10.0, trials=2, seed=9)      # 4 x 2 cells x 2 trials
    out = ev.run_sweep(agent, poses, record_trajectories=True)
    n = poses["start_x"].size
    assert out["reward_grid"].shape == (2, 4) and len(out["trajectory_endpoints"]) == n and len(out["trajectories"]) == 8
    oa = orc.OracleActor(sd)
    st0 = ev.start_states(poses["start_x"], poses["start_y"], poses["yaw_rad"], poses["L2"])
    for i in range(n):
        cfg = orc.default_cfg(); cfg.L2 = float(poses["L2"][i])
        o = orc.OracleEnv(cfg)
        obs = o.set_state(st0[i], (poses["start_x"][i], poses["start_y"][i], poses["yaw_rad"][i]))
        score, done, steps = 0.0, False, 0
        while not done:
            a = np.clip(oa.forward(obs[None])[0], -1, 1) * np.float32(0.78539819)
            obs, comps, done, viol, flags, succ = o.step(np.float32(a))
            score += comps[0]; steps += 1
        cls = ev.classify(np.array([int((flags * (1 << np.arange(6))).sum())], np.uint8), np.array([succ]))[0]
        ep = out["trajectory_endpoints"][i]
        assert ep["violation_type"] == ev.TERMINATION_CLASSES[cls], i
        assert abs(ep["score"] - score) < 2e-3 * max(1.0, abs(score)), (i, ep["score"], score)
        assert abs(ep["end_x"] - o.state[4]) < 1e-3 and abs(ep["end_y"] - o.state[5]) < 1e-3
    assert np.allclose(out["reward_grid"], out["trials"]["score"].reshape(2, 4, 2).mean(2))
    t0 = out["trajectories"][0]
    assert t0["trailer_x"][0] == pytest.approx(poses["start_x"][0]) and len(t0["trailer_x"]) >= 2


def test_episode_recorder_roundtrip(tmp_path):
    """f3: episodes recorded from the batched rollout replay exactly through the env (states, rewards), and the files are
    in the reference's formats."""
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import recording as rec
    N, track = 256, [0, 7, 100]
    env = tt.VecTruckTrailerEnv(N, seed=3, emit_info=True)
    agent = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=1 << 12, precision="fp32", actor_seed=0)
    agent.noise.bind_env(env)
    r = rec.EpisodeRecorder(env, track)
    obs, _ = env.reset()
    r.begin(obs)
    for it in range(400):
        raw = agent.choose_action(obs)
        scaled = agent.scale_action(raw)
        obs2, rew, done, info = env.step(scaled)
        r.record(raw, scaled, obs2, rew, done, info)
        obs, _ = env.reset(options={"mask": done})
        agent.noise.reset(done)
        r.after_reset(obs, done)
        if len(r.episodes) >= 4:
            break
    assert len(r.episodes) >= 4
    paths = r.save(str(tmp_path / "episode_replays"), str(tmp_path / "replay_buffer"))
    ep = r.episodes[0]
    d = pickle.load(open(paths[0], "rb"))
    T = len(d["actions"])
    assert len(d["states"]) == T + 1 and len(d["info"]) == T and d["info"][-1]["violation_type"] in tt.VIOLATION_NAMES
    assert os.path.basename(paths[0]) == f"episode_{ep['episode_num']}_reward_{int(sum(float(i['total_reward']) for i in d['info']))}.pkl"
    # replay the recorded actions from the recorded start state: same states and rewards
    env1 = tt.VecTruckTrailerEnv(1, emit_info=True)
    ed = d["env_data"]
    env1.set_state(d["states"][0].astype(np.float64)[None], [[ed["startx"], ed["starty"], ed["startyaw"]]], [[ed["goalx"], ed["goaly"], ed["goalyaw"]]])
    for t in range(T):
        _, rew, done, _ = env1.step(torch.from_numpy(d["actions"][t]).cuda())
        assert np.abs(env1.state[0].cpu().numpy().astype(np.float32) - d["states"][t + 1]).max() < 1e-5
        assert abs(float(rew[0]) - float(d["info"][t]["total_reward"])) < 1e-4 * max(1.0, abs(float(rew[0])))
    assert bool(done[0])
    tr = rec.load_transitions(paths[-1])
    assert len(tr) == len(r.episodes) and len(tr[0]) == T and tr[0][-1][4] is True and tr[0][0][0].shape == (23,)
    assert np.array_equal(tr[0][1][0], tr[0][0][3])                    # s_{t+1} of one transition is s of the next
    # reload into a replay buffer the reference way (trainv2.py:457-466)
    ag1 = tt.Agent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=4096)
    assert rec.remember_transitions(ag1, tr) == ag1.memory.mem_cntr


def test_save_and_load_models(tmp_path):
    """f4: Agent.save_models / load_models in the reference's file layout; the loaded policy is bit-identical."""
    import ddpg_trucktrailer_b200 as tt
    a = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=8, max_size=64, actor_seed=5)
    a.chkpt_dir = str(tmp_path / "tmp" / "ddpg")
    paths = a.save_models()
    assert set(paths) == {"actor", "target_actor"} and os.path.exists(os.path.join(a.chkpt_dir, "actor_ddpg"))
    a.save_models_progress(80)
    assert os.path.exists(os.path.join(a.chkpt_dir, "80", "actor_ddpg"))
    b = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=8, max_size=64, actor_seed=6)
    b.chkpt_dir = a.chkpt_dir
    obs = torch.empty(1000, 23, device="cuda").uniform_(-1, 1)
    assert not torch.equal(a.actor.forward(obs).clone(), b.actor.forward(obs).clone())
    assert "actor" in b.load_models()
    for prec in ("fp32", "f16"):
        assert torch.equal(a.actor.forward(obs, precision=prec).clone(), b.actor.forward(obs, precision=prec).clone())
    sd = torch.load(os.path.join(a.chkpt_dir, "actor_ddpg"))
    assert list(sd) == list(tt.ACTOR_KEYS) or set(sd) == set(tt.ACTOR_KEYS)
