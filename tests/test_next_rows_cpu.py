"""Host-side logic of the "next" rows of SURVEY.md section 8 (f2 evaluation sweep, f3 recording formats, f4 checkpoint
interop) -- no GPU needed.  Where /root/reference is mounted the files are also exchanged with the untouched reference."""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "ddpg-trucktrailer_b200")


def _load(name):
    """Import a pure-python module of the package without importing the package (whose __init__ is GPU-agnostic, but keep
    these tests independent of the CUDA library being built)."""
    import ddpg_trucktrailer_b200  # noqa: F401  (import shim; does not load the .so)
    return importlib.import_module("ddpg_trucktrailer_b200." + name)


def _actor_sd(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w0/")}, g


def test_checkpoint_files_and_roundtrip(tmp_path, golden_dir):
    ck = _load("checkpoint")
    sd, _ = _actor_sd(golden_dir)
    d = str(tmp_path / "tmp" / "ddpg")
    paths = ck.save_models({"actor": sd, "target_actor": sd}, d)
    assert paths["actor"] == os.path.join(d, "actor_ddpg") and os.path.exists(os.path.join(d, "target_actor_ddpg"))   # networks.py:19,108
    assert ck.save_models({"actor": sd}, d, progress=70)["actor"] == os.path.join(d, "70", "actor_ddpg")             # networks.py:78-82
    assert ck.checkpoint_path(d, "critic", best=True) == os.path.join(d, "critic_best")                               # networks.py:94
    back = ck.load_models(d)
    assert sorted(back) == ["actor", "target_actor"]
    assert all(torch.equal(back["actor"][k], sd[k]) for k in ck.ACTOR_KEYS)
    assert ck.check_actor_state_dict(back["actor"])
    bad = dict(sd); bad["fc1.weight"] = sd["fc1.weight"][:, :22]
    with pytest.raises(ValueError):
        ck.check_actor_state_dict(bad)
    with pytest.raises(FileNotFoundError):
        ck.load_models(d, names=("critic",), missing_ok=False)


def test_checkpoint_exchanged_with_reference(tmp_path, golden_dir):
    """Our file loads into the reference's ActorNetwork (networks.py:165-167) and reproduces its golden outputs; a file
    written by the reference's save_checkpoint (networks.py:149-152) loads here."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference not mounted")
    ck = _load("checkpoint")
    sd, g = _actor_sd(golden_dir)
    d = str(tmp_path / "ddpg")
    ck.save_state_dict(sd, d, "actor")
    actor = rh.make_actor(seed=123)                      # different initial weights
    actor.checkpoint_dir, actor.checkpoint_file = d, os.path.join(d, "actor_ddpg")
    actor.load_checkpoint()
    with torch.no_grad():
        out = actor.forward(torch.from_numpy(g["obs"])).numpy().reshape(-1)
    assert np.abs(out - g["out0"]).max() < 1e-6
    actor.checkpoint_file = os.path.join(d, "written_by_reference_ddpg")
    actor.save_checkpoint()
    back = ck.load_state_dict(d, "written_by_reference")
    assert ck.check_actor_state_dict(back) and all(torch.equal(back[k], sd[k]) for k in ck.ACTOR_KEYS)


def _fake_episode(T=7):
    rec = _load("recording")
    rng = np.random.default_rng(0)
    states = [rng.normal(size=6) for _ in range(T + 1)]
    actions = [rng.uniform(-0.7, 0.7, 1) for _ in range(T)]
    info = [rec.step_info_dict(10.4 + t, rng.normal(size=10), t % 8, t == T - 1) for t in range(T)]
    env_data = {"startx": 1.0, "starty": 2.0, "startyaw": 1.5, "goalx": 0, "goaly": -30, "goalyaw": np.pi / 2}
    return rec, states, actions, info, env_data


def test_episode_file_format(tmp_path):
    """episode_replay_collectorv2.py:20-33: file name, dict keys, element types."""
    rec, states, actions, info, env_data = _fake_episode()
    path = rec.save_episode(str(tmp_path), 42, states, actions, info, env_data)
    total = sum(10.4 + t for t in range(7))
    assert os.path.basename(path) == f"episode_42_reward_{int(total)}.pkl"
    d = pickle.load(open(path, "rb"))
    assert set(d) == {"states", "actions", "episode_num", "env_data", "info"} and d["episode_num"] == 42
    assert len(d["states"]) == 8 and d["states"][0].dtype == np.float32 and d["states"][0].shape == (6,)
    assert len(d["actions"]) == 7 and d["actions"][0].dtype == np.float32 and d["actions"][0].shape == (1,)
    assert d["info"][3]["violation_type"] == "major_boundary" and d["info"][6]["success"]
    assert sum(i["total_reward"] for i in d["info"]) == pytest.approx(total)                 # what save_episode sums (:29)
    assert d["env_data"]["goaly"] == -30


def test_episode_file_matches_reference_recording(tmp_path):
    """Same keys / element types as the reference's own recorded episode (DDPG/episode_replays/...pkl)."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference not mounted")
    ref = pickle.load(open(rh.golden_pkl_path(), "rb"))
    rec, states, actions, info, env_data = _fake_episode()
    mine = pickle.load(open(rec.save_episode(str(tmp_path), 1, states, actions, info, env_data), "rb"))
    assert set(mine) == set(ref) and set(mine["env_data"]) == set(ref["env_data"])
    assert type(mine["states"]) is type(ref["states"]) and mine["states"][0].dtype == ref["states"][0].dtype and mine["states"][0].shape == ref["states"][0].shape
    assert mine["actions"][0].dtype == ref["actions"][0].dtype and mine["actions"][0].shape == ref["actions"][0].shape
    consumed = {"total_reward", "distance_reward", "progress_reward", "heading_reward", "orientation_reward", "staged_success", "safety_penalty",
                "exploration_bonus", "final_success_bonus", "violation_type", "backward_penalty", "smoothness_penalty", "backward_movement_info", "success"}
    assert consumed <= set(mine["info"][0]) and set(ref["info"][0]) == consumed


def test_transitions_file_and_reload(tmp_path):
    """trainv2.py:333-351,457-466."""
    rec = _load("recording")
    rng = np.random.default_rng(1)
    eps = [[(rng.normal(size=23).astype(np.float32), rng.normal(size=1).astype(np.float32), float(t), rng.normal(size=23).astype(np.float32), t == 4)
            for t in range(5)] for _ in range(3)]
    path = rec.save_transitions(200, eps, str(tmp_path / "replay_buffer"))
    assert os.path.basename(path) == "transitions_episode_200_replay_buffer.pkl"
    back = rec.load_transitions(path)
    assert len(back) == 3 and len(back[0]) == 5 and back[2][4][4] is True

    class FakeAgent:
        def __init__(self): self.mem = []
        def remember(self, *t): self.mem.append(t)
    ag = FakeAgent()
    assert rec.remember_transitions(ag, back) == 15 and np.array_equal(ag.mem[7][0], eps[1][2][0])


def test_sweep_poses_and_classification():
    """heatmap.py:52-53,77-89,113-119,158-172."""
    ev = _load("evaluate")
    p = ev.heatmap_poses(seed=3)
    assert len(p["x_coords"]) == 30 and len(p["y_coords"]) == 15 and p["start_x"].size == 30 * 15 * 3
    assert p["yaw_deg"].min() >= 60 and p["yaw_deg"].max() <= 120 and p["L2"].min() >= 5 and p["L2"].max() <= 7
    assert p["start_x"][:3].tolist() == [-30.0] * 3 and p["start_y"][3 * 30] == 2.0            # trials of a cell are adjacent, rows = y
    st = ev.start_states(p["start_x"], p["start_y"], p["yaw_rad"], p["L2"])
    i = 77
    want = np.array([p["yaw_rad"][i], p["yaw_rad"][i], p["start_x"][i] + p["L2"][i] * np.cos(p["yaw_rad"][i]),
                     p["start_y"][i] + p["L2"][i] * np.sin(p["yaw_rad"][i]), p["start_x"][i], p["start_y"][i]], np.float32)
    assert np.array_equal(st[i], want.astype(np.float64))
    flags = np.array([0, 1, 2, 16, 4, 32, 1 | 2, 8, 16 | 4], np.uint8)
    succ = np.array([1, 0, 0, 0, 0, 0, 0, 0, 0], bool)
    names = [ev.TERMINATION_CLASSES[c] for c in ev.classify(flags, succ)]
    assert names == ["success", "jackknife", "out_of_map", "goal_passed", "max_steps", "other_failure", "jackknife", "success", "goal_passed"]


def test_learner_step_matches_reference_learn(golden_dir):
    """f1 checker: oracle/torch_learner.py (the torch restatement the CUDA learner is tested against on the GPU,
    tests/test_gpu_learner.py) == the reference's Agent.learn (DDPG_agent.py:72-131: critic MSE on the bootstrapped target
    with terminal masking, Adam (critic weight_decay 0.01), actor ascent on Q, soft target update) -- three consecutive
    steps from the reference's initial weights on the reference's batches (tests/golden/ref_learn.npz, generated by
    oracle/make_golden_learn.py from the untouched reference)."""
    from oracle import torch_learner as ln_mod
    g = np.load(os.path.join(golden_dir, "ref_learn.npz"))
    dims = tuple(int(v) for v in g["dims"])
    alpha, beta, tau, gamma = (float(v) for v in g["hyper"])
    B, steps = int(g["batch"]), int(g["steps"])
    sd = lambda phase, net: {k.split("/", 2)[2]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{phase}/{net}/")}

    class FakeActor:
        def __init__(self): self.dims, self._sd = dims, sd("before", "actor")
        def state_dict(self): return {k: v.clone() for k, v in self._sd.items()}
        def load_state_dict(self, s): self._sd = {k: v.clone() for k, v in s.items()}

    class FakeMemory:
        mem_cntr = B * steps
        def __init__(self): self.i = 0
        def sample_buffer(self, bs):
            j = slice(self.i * bs, (self.i + 1) * bs); self.i += 1
            t = torch.from_numpy
            return t(g["s"][j]), t(g["a"][j]), t(g["r"][j]), t(g["s2"][j]), t(g["d"][j])

    class FakeAgent:
        device = torch.device("cpu")
        def __init__(self):
            self.actor, self.memory = FakeActor(), FakeMemory()
            self.alpha, self.beta, self.tau, self.gamma, self.batch_size = alpha, beta, tau, gamma, B

    ag = FakeAgent()
    ln = ln_mod.TorchLearner(ag)
    for name in ("actor", "target_actor", "critic", "target_critic"):
        getattr(ln, name).load_state_dict(sd("before", name))
    for _ in range(steps):
        ln.learn()
    for name in ("actor", "target_actor", "critic", "target_critic"):
        want = sd("after", name)
        got = getattr(ln, name).state_dict()
        for k in want:
            assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-7), (name, k, (got[k] - want[k]).abs().max())
        moved = max((sd("before", name)[k] - want[k]).abs().max().item() for k in want)
        assert moved > (1e-7 if name.startswith("target") else 1e-5)   # the step really changed this network (targets: tau = 1e-3)
    # the updated policy was handed back to the rollout actor
    assert all(torch.allclose(ag.actor._sd[k], sd("after", "actor")[k], rtol=1e-5, atol=1e-7) for k in ag.actor._sd)
