"""Which output stream limits the env kernel?  K=1 launches with output arrays switched on/off (CUDA events)."""
import os, sys, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddpg_trucktrailer_b200 as tt
from ddpg_trucktrailer_b200 import _lib
N = 1 << 22
env = tt.VecTruckTrailerEnv(N, seed=27); env.reset()
L = tt.load(); s = _lib.stream_ptr()
act = torch.empty(1, N, device="cuda").uniform_(-0.3, 0.3)
obs = torch.empty(N, 23, device="cuda"); rew = torch.empty(N, device="cuda"); done = torch.empty(N, dtype=torch.uint8, device="cuda")
def run(o, r, d, iters=20):
    f = lambda: _lib.check(L.tt_env_step_k(env._h, act.data_ptr(), 1, 0, obs.data_ptr() if o else None, 23, rew.data_ptr() if r else None, done.data_ptr() if d else None, None, s))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    env.reset()
    return e0.elapsed_time(e1) / iters * 1e3
print("MINBLOCKS", os.environ.get("TT_ENV_MINBLOCKS", "default"), " obs+rew+done %.1f us | rew+done %.1f | none %.1f | obs only %.1f" % (run(1,1,1), run(0,1,1), run(0,0,0), run(1,0,0)))
