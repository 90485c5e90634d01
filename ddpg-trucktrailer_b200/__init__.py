"""B200-native batched truck-trailer rollout path (env step + DDPG actor rollout + replay store).

Host-side mirror of the reference's Python interface for this path (pain7576/ddpg-trucktrailer):
``Truck_trailer_Env_2.reset/step`` (truck_trailer_sim/simv2.py), ``Agent.choose_action/remember``
(DDPG/DDPG_agent.py), ``ReplayBuffer.store_transition/sample_buffer`` (DDPG/replay_buffer.py) -- same names,
argument meaning and return conventions, with a leading ``N`` (environment) dimension.  All compute is in
hand-written sm_100a CUDA behind the C ABI of ``include/tt_b200.h``; PyTorch only owns device memory,
streams and ``torch.distributed``.  There is no CPU fallback.
"""
from ._lib import TTError, load, lib_path, COMP_NAMES, VIOLATION_NAMES, FLAG_NAMES, STAT_NAMES  # noqa: F401
from .env import VecTruckTrailerEnv, Truck_trailer_Env_2, EnvConfig  # noqa: F401
from .replay import DeviceReplayBuffer, ReplayBuffer  # noqa: F401
from .agent import VecAgent, Agent, OUNoiseState, init_actor_state_dict, ACTOR_KEYS  # noqa: F401
from .rollout import RolloutEngine, AsyncTrainer  # noqa: F401
from . import checkpoint, recording, evaluate  # noqa: F401

__all__ = ["VecTruckTrailerEnv", "Truck_trailer_Env_2", "EnvConfig", "DeviceReplayBuffer", "ReplayBuffer", "VecAgent",
           "Agent", "OUNoiseState", "RolloutEngine", "AsyncTrainer", "init_actor_state_dict", "TTError", "load", "lib_path"]
