"""Latency of the N = 1 drop-in wrappers (the reference's single-env contract, DDPG/trainv2.py:488-531) -- wall clock per call,
host round trips included -- next to the device-side cost of the same calls.
    python profiles/n1_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ddpg_trucktrailer_b200 as tt

env = tt.Truck_trailer_Env_2(seed=27)
agent = tt.Agent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=4096)
obs, _ = env.reset(seed=27)


def wall(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

state = {"obs": obs, "ep": 0}
def loop_body():
    a = agent.choose_action(state["obs"])
    scaled = np.clip(a, -1, 1) * env.action_space.high
    o2, r, d, info = env.step(scaled)
    agent.remember(state["obs"], a, r, o2, d)
    state["obs"] = o2
    if d:
        state["ep"] += 1
        state["obs"], _ = env.reset(seed=27 + state["ep"]); agent.noise.reset()
print(f"Agent.choose_action (numpy in, numpy out):            {wall(lambda: agent.choose_action(state['obs'])):8.1f} us per call  (precision {agent.precision} -> {agent.actor.auto_precision(1)})")
a = np.array([0.2], np.float32)
def step_only():
    o2, r, d, info = env.step(a)
    if d: env.reset(seed=1)
print(f"Truck_trailer_Env_2.step (python floats / info dict out):   {wall(step_only):8.1f} us per call")
print(f"whole loop body choose_action + step + remember:            {wall(loop_body):8.1f} us per iteration")
# device side of the same work, no host round trip: the vector classes at N = 1
venv = tt.VecTruckTrailerEnv(1, seed=27); vag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=1, max_size=4096)
eng = tt.RolloutEngine(venv, vag); eng.reset()
print(f"RolloutEngine.step at N = 1 (2 launches, no host sync):    {wall(eng.step, 2000):8.1f} us per iteration")
