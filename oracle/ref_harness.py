"""TEST INFRASTRUCTURE ONLY -- loader for the *untouched* reference (pain7576/ddpg-trucktrailer).

Imports the reference's own modules from ``/root/reference`` (read-only, present only in the build
container, never on the GPU box) so that ``oracle/make_golden.py`` can generate golden vectors and
``tests/test_oracle_vs_reference.py`` can pin the C restatement in ``oracle/tt_oracle.c`` against it.

Nothing in the product package (``ddpg-trucktrailer_b200/``), in ``bench.py`` or in the ``-m gpu``
tests may import this file.

The reference needs three things that are absent in this image:
  * ``gym`` / ``gym.spaces.Box`` / ``gym.error``     -> minimal stand-ins (only ``.low/.high/.shape``)
  * ``matplotlib`` (+ pyplot / patches / transforms) -> empty modules, used only by ``render()``
  * a CUDA device for ``networks.py:51,134``         -> ``nn.Module.to`` patched to a no-op while
                                                        the networks are constructed
No reference source is modified or copied.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "truck_trailer_sim", "simv2.py"))


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # gym.Env stand-in (simv2.py:20)
            reward_range = (-float("inf"), float("inf"))
            metadata: dict = {}

        class Box:  # gym.spaces.Box stand-in (simv2.py:79-91)
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                self.shape = tuple(shape) if shape is not None else np.shape(low)
                self.low = np.full(self.shape, low, dtype=self.dtype)
                self.high = np.full(self.shape, high, dtype=self.dtype)

        spaces = types.ModuleType("gym.spaces")
        spaces.Box = Box
        error = types.ModuleType("gym.error")
        gym.Env, gym.spaces, gym.error = Env, spaces, error
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.error": error})

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        patches = types.ModuleType("matplotlib.patches")
        transforms = types.ModuleType("matplotlib.transforms")
        for name in ("Rectangle", "Circle", "FancyArrow"):
            setattr(patches, name, type(name, (), {}))
        transforms.Affine2D = type("Affine2D", (), {})
        mpl.pyplot, mpl.patches, mpl.transforms = plt, patches, transforms
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt,
                            "matplotlib.patches": patches, "matplotlib.transforms": transforms})


def load_env_module():
    """Return the reference module ``truck_trailer_sim.simv2`` (class ``Truck_trailer_Env_2``)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module("truck_trailer_sim.simv2")


def make_env():
    return load_env_module().Truck_trailer_Env_2()


@contextlib.contextmanager
def _module_to_is_noop():
    import torch.nn as nn
    orig = nn.Module.to
    nn.Module.to = lambda self, *a, **k: self  # networks.py:51,134 hard-code a CUDA device
    try:
        yield
    finally:
        nn.Module.to = orig


def load_ddpg_modules():
    """Return (networks, noise, replay_buffer) reference modules from ``DDPG/``."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    ddpg = os.path.join(REFERENCE_ROOT, "DDPG")
    if ddpg not in sys.path:
        sys.path.insert(0, ddpg)
    import importlib
    return (importlib.import_module("networks"), importlib.import_module("noise"),
            importlib.import_module("replay_buffer"))


def make_actor(seed: int = 0, input_dims=(23,), fc1=400, fc2=300, n_actions=1):
    """Reference ``ActorNetwork`` on CPU with ``torch.manual_seed(seed)`` initial weights."""
    import torch
    networks, _, _ = load_ddpg_modules()
    torch.manual_seed(seed)
    with _module_to_is_noop():
        actor = networks.ActorNetwork(1e-4, list(input_dims), fc1, fc2, n_actions=n_actions, name="actor")
    actor.device = torch.device("cpu")
    actor.eval()
    return actor


def golden_pkl_path() -> str:
    return os.path.join(REFERENCE_ROOT, "DDPG", "episode_replays", "episode_10579_reward_4792.pkl")


REWARD_KEYS = ("total_reward", "distance_reward", "progress_reward", "heading_reward",
               "orientation_reward", "staged_success", "safety_penalty", "exploration_bonus",
               "final_success_bonus", "backward_penalty", "smoothness_penalty")

VIOLATION_CODES = {"none": 0, "jackknife": 1, "jackknife_warning": 2, "major_boundary": 3,
                   "minor_boundary": 4, "past_the_goal": 5, "max_step": 6, "excessive_backward": 7}


def rollout_reference(env, state0, start, goal, actions, stop_on_done=True):
    """Replay ``actions`` (scaled steering, float32) through the untouched reference from an
    injected pose, the way ``DDPG/test.py:96-115`` and ``episode_replay_collectorv2.py:258`` do.

    Returns dict of per-step arrays: state f64[T,6], obs f32[T,23], comps f64[T,11], viol u8[T],
    flags u8[T,6] (jackknife,out_of_map,max_steps,goal_reached,goal_passed,excessive_backward),
    done u8[T], success u8[T]; plus obs0 f32[23], max_steps.
    """
    env.reset(seed=0)
    env.startx, env.starty, env.startyaw = (float(v) for v in start)
    env.goalx, env.goaly, env.goalyaw = goal
    env.state = np.asarray(state0).copy()
    env.max_episode_steps = env.compute_max_steps()
    env.episode_steps = 0
    env.reward_state = None
    obs0 = env.compute_observation(env.state, 0.0)
    T = len(actions)
    out = dict(state=np.zeros((T, 6)), obs=np.zeros((T, 23), np.float32), comps=np.zeros((T, 11)),
               viol=np.zeros(T, np.uint8), flags=np.zeros((T, 6), np.uint8), done=np.zeros(T, np.uint8),
               success=np.zeros(T, np.uint8))
    n = 0
    for t in range(T):
        a = np.asarray(actions[t], dtype=np.float32).reshape(1)
        obs, rew, done, info = env.step(a)
        out["state"][t] = env.state
        out["obs"][t] = obs
        out["comps"][t] = [float(info[k]) for k in REWARD_KEYS]
        out["viol"][t] = VIOLATION_CODES[info["violation_type"]]
        rs = env.reward_state
        exb = float(np.hypot(env.state[4] - env.goalx, env.state[5] - env.goaly)) > rs["closest_distance_to_goal"] + 6.0
        out["flags"][t] = [env.jackknife, env.out_of_map, env.max_steps_reached, env.goal_reached,
                           env.goal_passed, exb]
        out["done"][t] = bool(done)
        out["success"][t] = bool(info["success"])
        n = t + 1
        if done and stop_on_done:
            break
    for k in ("state", "obs", "comps", "viol", "flags", "done", "success"):
        out[k] = out[k][:n]
    out["obs0"] = obs0
    out["max_steps"] = env.max_episode_steps
    return out
