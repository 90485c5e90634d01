"""Episode / transition recording in the reference's on-disk formats.

* ``episode_<n>_reward_<R>.pkl`` -- ``EpisodeReplaySystem.save_episode`` (DDPG/episode_replay_collectorv2.py:20-33,
  same in episode_replay_collector.py): a dict ``{'states', 'actions', 'episode_num', 'env_data', 'info'}`` with
  ``states`` = list of T+1 float32 (6,) arrays (``env.state.copy()`` before the first and after every step,
  trainv2.py:499,522), ``actions`` = list of T float32 (1,) scaled steering arrays (trainv2.py:516-518), ``info`` = list
  of T reward-info dicts (reward_functionv1.py:488-504), ``env_data`` = start / goal pose (trainv2.py:502-509);
  ``R`` = int(sum of total_reward).
* ``transitions_episode_<n>_replay_buffer.pkl`` -- ``save_transitions`` (DDPG/trainv2.py:333-339): a list (episodes) of
  lists of ``(obs, action, reward, obs_next, done)`` tuples, re-loaded with ``agent.remember`` (trainv2.py:457-466).

``EpisodeRecorder`` collects these from the batched CUDA rollout for a chosen set of environments: per step it copies the
tracked rows of the device tensors to the host (a few hundred bytes per tracked env) and cuts episodes at ``done``.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from ._lib import COMP_NAMES, VIOLATION_NAMES

INFO_KEYS = ("total_reward",) + tuple(COMP_NAMES)


def step_info_dict(total, comps, violation_code, success):
    """One reward-info dict with the reference's keys (reward_functionv1.py:488-504).  ``backward_movement_info`` is
    reduced to the penalty (its other entries are diagnostics no consumer in the reference reads)."""
    d = {"total_reward": np.float64(total)}
    for k, v in zip(COMP_NAMES, comps):
        d[k] = np.float64(v)
    d["violation_type"] = VIOLATION_NAMES[int(violation_code)]
    d["backward_movement_info"] = {"penalty": float(d.get("backward_penalty", 0.0))}
    d["success"] = np.bool_(bool(success))
    return d


def save_episode(save_dir, episode_num, states, actions, info, env_data=None):
    """``EpisodeReplaySystem.save_episode`` (episode_replay_collectorv2.py:20-33); returns the file path."""
    os.makedirs(save_dir, exist_ok=True)
    data = {"states": [np.asarray(s, np.float32).reshape(6) for s in states],
            "actions": [np.asarray(a, np.float32).reshape(1) for a in actions],
            "episode_num": episode_num, "env_data": env_data, "info": list(info)}
    total = sum(float(i["total_reward"]) for i in info)
    path = os.path.join(save_dir, f"episode_{episode_num}_reward_{int(total)}.pkl")
    with open(path, "wb") as f:
        pickle.dump(data, f)
    return path


def load_episode(path):
    """``load_episode`` of the replay tools (episode_replay_collectorv2.py / episode_playerv2.py)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def save_transitions(episode_num, transitions_history, save_dir="replay_buffer"):
    """``save_transitions`` (trainv2.py:333-339)."""
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, f"transitions_episode_{episode_num}_replay_buffer.pkl")
    with open(path, "wb") as f:
        pickle.dump(list(transitions_history), f)
    return path


def load_transitions(path):
    """``load_transitions`` (trainv2.py:341-351) -> list of episodes of (obs, action, reward, obs_next, done)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def remember_transitions(agent, transitions):
    """trainv2.py:457-466: feed stored transitions to ``agent.remember`` (any object with the reference signature)."""
    n = 0
    for ep in transitions:
        for obs, action, reward, obs_next, done in ep:
            agent.remember(obs, action, reward, obs_next, done)
            n += 1
    return n


class EpisodeRecorder:
    """Records whole episodes of the environments ``track`` (indices into the batch) from a stepping loop:

        rec = EpisodeRecorder(env, track=[0, 5, 9])          # env: VecTruckTrailerEnv(..., emit_info=True)
        obs, _ = env.reset()
        rec.begin(obs)
        while ...:
            raw = agent.choose_action(obs); scaled = agent.scale_action(raw)
            obs2, rew, done, info = env.step(scaled)
            rec.record(raw, scaled, obs2, rew, done, info)    # BEFORE the masked reset
            obs, _ = env.reset(options={'mask': done})
            rec.after_reset(obs, done)

    Finished episodes accumulate in ``rec.episodes`` (dicts in the save_episode layout plus ``transitions``) and can be
    written with ``rec.save(dir)``.
    """

    def __init__(self, env, track):
        import torch
        self.env = env
        self.track = torch.as_tensor(list(track), dtype=torch.int64, device=env.device)
        self.n = len(self.track)
        self.episodes = []
        self._open = [None] * self.n
        self._count = 0

    def _pose(self):
        s = self.env.get_state()
        return (s["state"][self.track].cpu().numpy(), s["start"][self.track].cpu().numpy(), s["goal"][self.track].cpu().numpy())

    def _start(self, j, obs_row, state, start, goal):
        self._open[j] = dict(states=[state.astype(np.float32)], actions=[], info=[], transitions=[], obs=obs_row.copy(),
                             env_data={"startx": float(start[0]), "starty": float(start[1]), "startyaw": float(start[2]),
                                       "goalx": float(goal[0]), "goaly": float(goal[1]), "goalyaw": float(goal[2])})

    def begin(self, obs):
        st, sp, gl = self._pose()
        o = obs[self.track].cpu().numpy()
        for j in range(self.n):
            self._start(j, o[j], st[j], sp[j], gl[j])

    def record(self, raw_action, scaled_action, obs_next, reward, done, info):
        if not info:
            raise ValueError("EpisodeRecorder needs the env's info (emit_info=True)")
        t = self.track
        raw = raw_action.reshape(-1)[t].cpu().numpy(); sc = scaled_action.reshape(-1)[t].cpu().numpy()
        o2 = obs_next[t].cpu().numpy(); r = reward[t].cpu().numpy(); d = done[t].cpu().numpy().astype(bool)
        comps = np.stack([info[k][t].cpu().numpy() for k in COMP_NAMES], 1)
        viol = info["violation_type"][t].cpu().numpy(); succ = info["success"][t].cpu().numpy()
        st, _, _ = self._pose()
        for j in range(self.n):
            ep = self._open[j]
            if ep is None:
                continue
            ep["actions"].append(np.array([sc[j]], np.float32))
            ep["states"].append(st[j].astype(np.float32))
            ep["info"].append(step_info_dict(r[j], comps[j], viol[j], succ[j]))
            ep["transitions"].append((ep["obs"].copy(), np.array([raw[j]], np.float32), float(r[j]), o2[j].copy(), bool(d[j])))
            ep["obs"] = o2[j].copy()
            if d[j]:
                ep["episode_num"] = self._count
                self._count += 1
                del ep["obs"]
                self.episodes.append(ep)
                self._open[j] = None

    def after_reset(self, obs, done):
        d = done[self.track].cpu().numpy().astype(bool)
        if not d.any():
            return
        st, sp, gl = self._pose()
        o = obs[self.track].cpu().numpy()
        for j in np.nonzero(d)[0]:
            self._start(int(j), o[j], st[j], sp[j], gl[j])

    def save(self, save_dir, transitions_dir=None):
        paths = [save_episode(save_dir, ep["episode_num"], ep["states"], ep["actions"], ep["info"], ep["env_data"]) for ep in self.episodes]
        if transitions_dir is not None and self.episodes:
            paths.append(save_transitions(self.episodes[-1]["episode_num"], [ep["transitions"] for ep in self.episodes], transitions_dir))
        return paths
