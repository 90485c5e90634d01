"""Times the UNMODIFIED reference Python loop (baseline/_ref, see prepare_ref.py) on the host cores.

    mode "rollout": the body of DDPG/trainv2.py:488-531 without learn(): agent.choose_action(obs) -> clip(a, -1, 1) * high ->
                    env.step -> agent.remember, env.reset(seed) + agent.noise.reset() on done   (the metric's path)
    mode "env":     env.step with random steering U(-pi/4, pi/4) (float32) + env.reset(seed) on done   

Single process or one process per core (multiprocessing, OMP/MKL threads = 1).  The reference hard-codes a CUDA device for
its networks (networks.py:51,134); the workers hide the GPUs and construct the Agent under a no-op nn.Module.to so that the
baseline is the reference on the CPU, as BASELINE.md section 3 specifies.  Harness only: no reference file is modified."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")


def _setup_paths():
    for p in (os.path.join(REF, "DDPG"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name, mod in (("gym", "gym"), ("matplotlib", "matplotlib")):
        try:
            __import__(mod)
        except Exception:
            if STUBS not in sys.path:
                sys.path.append(STUBS)


def _make(mode, seed):
    import numpy as np
    _setup_paths()
    from truck_trailer_sim.simv2 import Truck_trailer_Env_2
    env = Truck_trailer_Env_2()
    agent = None
    if mode == "rollout":
        import torch
        import torch.nn as nn
        torch.set_num_threads(1)
        torch.manual_seed(seed)
        from DDPG_agent import Agent
        orig = nn.Module.to
        nn.Module.to = lambda self, *a, **k: self            # networks.py:51,134: T.device('cuda:0' ...) -> stay on the CPU
        try:
            agent = Agent(alpha=1e-4, beta=1e-3, input_dims=[23], tau=1e-3, n_actions=1, batch_size=64, max_size=200000)
        finally:
            nn.Module.to = orig
        for net in (agent.actor, agent.critic, agent.target_actor, agent.target_critic):
            net.device = torch.device("cpu")
    return env, agent, np


class Loop:
    """The reference loop as a resumable object: run(n) advances n env steps."""

    def __init__(self, mode, seed=27):
        self.mode, self.seed = mode, seed
        self.env, self.agent, np = _make(mode, seed)
        self.np = np
        self.rng = np.random.default_rng(seed)
        self.episode = 0
        self.obs, _ = self.env.reset(seed=seed)
        if self.agent is not None:
            self.agent.noise.reset()
        self.episodes_done = 0

    def run(self, n):
        np, env, agent = self.np, self.env, self.agent
        obs = self.obs
        for _ in range(n):
            if agent is not None:                                 # trainv2.py:512-526
                action = agent.choose_action(obs)
                scaled = np.clip(action, -1, 1) * env.action_space.high
                obs_, reward, done, info = env.step(scaled)
                agent.remember(obs, action, reward, obs_, done)
            else:
                a = self.rng.uniform(-np.pi / 4, np.pi / 4, 1).astype(np.float32)
                obs_, reward, done, info = env.step(a)
            obs = obs_
            if done:                                              # trainv2.py:489-492
                self.episode += 1; self.episodes_done += 1
                obs, _ = env.reset(seed=self.seed + self.episode)
                if agent is not None:
                    agent.noise.reset()
        self.obs = obs


def _worker(mode, seed, warm, chunks, chunk, barrier, out):
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import warnings
    warnings.filterwarnings("ignore")          # DDPG_agent.py:38 builds a tensor from a list of ndarrays on every call
    try:
        lp = Loop(mode, seed)
        lp.run(warm)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(chunks):
            lp.run(chunk)
        out.put((chunks * chunk, time.perf_counter() - t0, lp.episodes_done))
    except Exception as e:           # never leave the parent hanging on the barrier / queue
        try:
            barrier.abort()
        except Exception:
            pass
        out.put(("error", repr(e), 0))


def time_loop(mode, procs, steps_per_proc, warm=300, chunks=1):
    """Returns dict(value = aggregate env-steps/s over `procs` processes, steps, seconds, procs)."""
    ctx = mp.get_context("spawn")
    barrier, out = ctx.Barrier(procs), ctx.Queue()
    chunk = max(1, steps_per_proc // chunks)
    ps = [ctx.Process(target=_worker, args=(mode, 1000 + 97 * i, warm, chunks, chunk, barrier, out), daemon=True) for i in range(procs)]
    for p in ps:
        p.start()
    res = [out.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    bad = [r for r in res if r[0] == "error"]
    if bad:
        raise RuntimeError(f"reference loop failed: {bad[0][1]}")
    steps = sum(r[0] for r in res)
    secs = max(r[1] for r in res)
    return {"value": steps / secs, "steps": steps, "seconds": secs, "procs": procs, "episodes": sum(r[2] for r in res)}


if __name__ == "__main__":
    import json
    mode = sys.argv[1] if len(sys.argv) > 1 else "rollout"
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    print(json.dumps(time_loop(mode, procs, steps)))
