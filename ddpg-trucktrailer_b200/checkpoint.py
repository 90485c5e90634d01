"""Checkpoint interop with the reference.

The reference stores each of its four networks as ``torch.save(module.state_dict(), '<chkpt_dir>/<name>_ddpg')`` with
``name`` in ``actor, target_actor, critic, target_critic`` (DDPG/networks.py:10-19,69-95,99-108,149-169); progress
checkpoints go to ``<chkpt_dir>/<success>/<name>_ddpg``, the best model to ``<chkpt_dir>/<name>_best``.  These helpers
read and write exactly those files, so that a policy trained by the reference can be rolled out by the CUDA actor and
weights produced here can be loaded by the reference's ``test.py`` / ``heatmap.py`` unchanged.
"""
from __future__ import annotations

import os

import torch

NETWORK_NAMES = ("actor", "target_actor", "critic", "target_critic")
ACTOR_KEYS = ("fc1.weight", "fc1.bias", "bn1.weight", "bn1.bias", "fc2.weight", "fc2.bias", "bn2.weight", "bn2.bias",
              "mu.weight", "mu.bias")
CRITIC_KEYS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias",
               "action_value.weight", "action_value.bias", "q.weight", "q.bias")


def checkpoint_path(chkpt_dir, name, progress=None, best=False):
    """networks.py:19 / :78-82 / :94 file naming."""
    if best:
        return os.path.join(chkpt_dir, name + "_best")
    if progress is not None:
        return os.path.join(chkpt_dir, str(progress), name + "_ddpg")
    return os.path.join(chkpt_dir, name + "_ddpg")


def _cpu(sd):
    return {k: torch.as_tensor(v).detach().to("cpu").clone() for k, v in sd.items()}


def save_state_dict(sd, chkpt_dir, name, progress=None, best=False):
    """``T.save(self.state_dict(), file)`` of networks.py:69-95,149-169 (tensors are written from the CPU so the file
    loads on a machine without the GPU it was trained on)."""
    path = checkpoint_path(chkpt_dir, name, progress, best)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(_cpu(sd), path)
    return path


def load_state_dict(chkpt_dir, name, progress=None, best=False, map_location="cpu"):
    """``T.load(self.checkpoint_file)`` of networks.py:87-89,165-167."""
    sd = torch.load(checkpoint_path(chkpt_dir, name, progress, best), map_location=map_location)
    if not isinstance(sd, dict):
        raise ValueError("not a state_dict checkpoint")
    return sd


def check_actor_state_dict(sd, input_dims=23, fc1_dims=400, fc2_dims=300):
    """Raises ValueError unless ``sd`` has the reference ActorNetwork's keys and shapes (networks.py:110-131)."""
    shapes = {"fc1.weight": (fc1_dims, input_dims), "fc1.bias": (fc1_dims,), "bn1.weight": (fc1_dims,), "bn1.bias": (fc1_dims,),
              "fc2.weight": (fc2_dims, fc1_dims), "fc2.bias": (fc2_dims,), "bn2.weight": (fc2_dims,), "bn2.bias": (fc2_dims,),
              "mu.weight": (1, fc2_dims), "mu.bias": (1,)}
    for k, shp in shapes.items():
        if k not in sd:
            raise ValueError(f"actor checkpoint lacks {k}")
        if tuple(sd[k].shape) != shp:
            raise ValueError(f"{k}: expected {shp}, got {tuple(sd[k].shape)}")
    return True


def save_models(state_dicts, chkpt_dir="tmp/ddpg", progress=None):
    """``Agent.save_models`` / ``save_models_progress`` (DDPG_agent.py:54-64): ``state_dicts`` maps network name ->
    state_dict for any subset of NETWORK_NAMES."""
    return {name: save_state_dict(sd, chkpt_dir, name, progress) for name, sd in state_dicts.items()}


def load_models(chkpt_dir="tmp/ddpg", names=NETWORK_NAMES, progress=None, missing_ok=True):
    """``Agent.load_models`` (DDPG_agent.py:66-70) -> dict name -> state_dict (files that do not exist are skipped when
    ``missing_ok``; the actor is what the rollout path needs)."""
    out = {}
    for name in names:
        path = checkpoint_path(chkpt_dir, name, progress)
        if os.path.exists(path):
            out[name] = load_state_dict(chkpt_dir, name, progress)
        elif not missing_ok:
            raise FileNotFoundError(path)
    return out
