"""Max abs error of the tensor-core actor (both precisions) against the reference torch outputs in tests/golden/ref_actor.npz."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ddpg_trucktrailer_b200 as tt
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_agent import _sets
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_actor.npz"))
obs = torch.from_numpy(g["obs"]).cuda()
actor = tt.agent.CudaActor()
for name, w, ref in zip(("init", "amplified"), _sets(g), (g["out0"], g["out1"])):
    actor.load_state_dict(w)
    for prec in ("fp32", "f16", "f16_plain", "bf16"):
        out = actor.forward(obs, precision=prec).cpu().numpy()
        print(f"{name:10s} {prec:5s} max|err| = {np.abs(out - ref).max():.3e}")
