"""TEST INFRASTRUCTURE: golden vectors for the learner step.

Constructs the UNTOUCHED reference ``Agent`` (DDPG/DDPG_agent.py) on the CPU (the hard-coded CUDA device of
networks.py:51,134 is neutralised by a no-op ``nn.Module.to``), fills its replay buffer with a fixed set of transitions,
makes ``sample_buffer`` return them in order, runs ``learn()`` three times and stores the four networks before and
after.  Small hidden sizes (fc1 = 64, fc2 = 48 are constructor arguments of the reference Agent) keep the fixture small.
Output: tests/golden/ref_learn.npz.  Needs /root/reference (build container only).

    python oracle/make_golden_learn.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as rh      # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NETS = ("actor", "target_actor", "critic", "target_critic")


def main(seed=5, fc1=64, fc2=48, batch=64, steps=3):
    rh.load_ddpg_modules()
    import importlib
    ddpg_agent = importlib.import_module("DDPG_agent")
    torch.manual_seed(seed)
    with rh._module_to_is_noop():
        agent = ddpg_agent.Agent(alpha=1e-4, beta=1e-3, input_dims=[23], tau=1e-3, n_actions=1, fc1_dims=fc1, fc2_dims=fc2, batch_size=batch)
    for n in NETS:
        getattr(agent, n).device = torch.device("cpu")
    rng = np.random.default_rng(seed)
    B = batch * steps
    s = rng.uniform(-1, 1, (B, 23)).astype(np.float32); s2 = rng.uniform(-1, 1, (B, 23)).astype(np.float32)
    a = rng.uniform(-1.2, 1.2, (B, 1)).astype(np.float32); r = rng.normal(5, 30, B).astype(np.float32); d = rng.random(B) < 0.1
    for i in range(B):
        agent.remember(s[i], a[i], r[i], s2[i], d[i])
    order = iter(np.arange(B).reshape(steps, batch))
    mem = agent.memory
    mem.sample_buffer = lambda bs: (lambda idx: (mem.state_memory[idx], mem.action_memory[idx], mem.reward_memory[idx],
                                                 mem.new_state_memory[idx], mem.terminal_memory[idx]))(next(order))
    out = {}
    for n in NETS:
        for k, v in getattr(agent, n).state_dict().items():
            out[f"before/{n}/{k}"] = v.detach().numpy().copy()
    for _ in range(steps):
        agent.learn()
    for n in NETS:
        for k, v in getattr(agent, n).state_dict().items():
            out[f"after/{n}/{k}"] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ref_learn.npz"), s=s, a=a, r=r, s2=s2, d=d, dims=np.array([23, fc1, fc2]), steps=steps, batch=batch,
                        hyper=np.array([1e-4, 1e-3, 1e-3, 0.99]), **out)
    print("wrote ref_learn.npz", sum(v.size for v in out.values()), "parameters")


if __name__ == "__main__":
    main()
