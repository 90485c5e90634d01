"""The whole rollout iteration (trainv2.py:511-531 without learn) through tt_rollout_step vs the same
sequence driven by hand through the reference-shaped Python API, and vs the oracle on a small batch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(tt, N, cap, seed=11):
    env = tt.VecTruckTrailerEnv(N, seed=seed)
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=cap, actor_seed=0, seed=seed)
    sd = tt.init_actor_state_dict(seed=0); sd["mu.weight"] *= 100      # make the policy move the steering
    ag.load_actor_state_dict(sd)
    return env, ag


def test_rollout_step_equals_manual_loop():
    import ddpg_trucktrailer_b200 as tt
    N, cap, iters = 3000, 10000, 30
    env1, ag1 = _mk(tt, N, cap); eng = tt.RolloutEngine(env1, ag1); eng.reset()
    env2, ag2 = _mk(tt, N, cap); ag2.noise.bind_env(env2)
    obs, _ = env2.reset(); ag2.noise.reset()
    tot_done = 0
    for it in range(iters):
        o1, r1, d1 = eng.step()
        # reference-shaped loop body (trainv2.py:512-531)
        action = ag2.choose_action(obs)
        scaled = ag2.scale_action(action)
        obs_, reward, done, info = env2.step(scaled)
        ag2.remember(obs, action, reward, obs_, done)
        ag2.noise.reset(mask=done)
        obs, _ = env2.reset(options={"mask": done})
        env2.tick()
        assert torch.equal(r1, reward) and torch.equal(d1.bool(), done), it
        assert torch.equal(o1, obs), it
        tot_done += int(done.sum())
    assert tot_done > 0
    m1, m2 = ag1.memory, ag2.memory
    assert m1.mem_cntr == m2.mem_cntr == N * iters
    for f in ("state_memory", "new_state_memory", "action_memory", "reward_memory", "terminal_memory"):
        assert torch.equal(getattr(m1, f), getattr(m2, f)), f
    s1, s2 = env1.read_stats(), env2.read_stats()
    assert s1 == s2 and s1["steps"] == N * iters and s1["episodes"] == tot_done


def test_rollout_vs_oracle_small():
    """Closed loop (actor in the loop) against the float64 oracle on 64 envs, evaluate=True (no noise)."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    N = 64
    env, ag = _mk(tt, N, 4096, seed=3)
    eng = tt.RolloutEngine(env, ag, evaluate=True)
    obs = eng.reset().clone()
    s = env.get_state()
    sd = {k: v.cpu().numpy() for k, v in ag.actor.state_dict().items()}
    oa = orc.OracleActor(sd)
    envs = []
    for i in range(N):
        e = orc.OracleEnv(); e.reset_pose(*s["start"][i].cpu().numpy()); envs.append(e)
    alive = np.ones(N, bool)
    oobs = obs.cpu().numpy().copy()
    for it in range(60):
        o1, r1, d1 = eng.step()
        mu = oa.forward(oobs)
        for i in np.nonzero(alive)[0]:
            scaled = np.float32(np.clip(mu[i], -1, 1)) * np.float32(0.78539819)
            o, comps, done, viol, flags, succ = envs[i].step(scaled)
            assert abs(float(r1[i]) - comps[0]) < 1e-3 * max(1.0, abs(comps[0])), (it, i)
            assert bool(d1[i]) == done, (it, i)
            oobs[i] = o
            if done:
                alive[i] = False          # the CUDA env auto-resets; stop comparing this env
        if not alive.any():
            break
    assert (~alive).sum() > 0


@pytest.mark.parametrize("precision", ["fp32", "f16"])
@pytest.mark.parametrize("N,cap,cntr0", [(1000, 4096, 0), (3000, 10000, 9000), (5000, 2048, 77), (4096, 4096, 4096 * 3 + 4), (777, 1 << 16, 5)])
def test_fused_store_equals_scatter_kernel(precision, N, cap, cntr0):
    """The producers' fused ring writes (actor: s, noise kernel: a, env kernel: s', r, done) put exactly the bytes
    where the stand-alone scatter kernel / ReplayBuffer.store_transition semantics put them: wrap-around, batches
    larger than the ring (last writer wins), unaligned ring positions, ragged tiles."""
    import ctypes as C
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import _lib
    L = tt.load(); s = _lib.stream_ptr()
    env = tt.VecTruckTrailerEnv(N, seed=4); ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=cap, actor_seed=1, precision=precision)
    ag.noise.bind_env(env)
    obs, _ = env.reset()
    obs = obs.clone()
    ref = tt.DeviceReplayBuffer(cap); ref.mem_cntr = cntr0
    fused = tt.DeviceReplayBuffer(cap)
    for mem in (ref, fused):                                  # sentinel so that untouched rows are compared too
        mem.state_memory.fill_(-7.0); mem.new_state_memory.fill_(-7.0); mem.action_memory.fill_(-7.0); mem.reward_memory.fill_(-7.0); mem.terminal_memory.fill_(9)
    ring = _lib.ReplayRing(fused.state_memory.data_ptr(), fused.action_memory.data_ptr(), fused.reward_memory.data_ptr(),
                           fused.new_state_memory.data_ptr(), fused.terminal_memory.data_ptr(), cap, cntr0)
    mu = torch.empty(N, device="cuda"); scaled = torch.empty(N, device="cuda"); x = torch.zeros(N, device="cuda")
    obs2 = torch.empty(N, 23, device="cuda"); rew = torch.empty(N, device="cuda"); done = torch.empty(N, dtype=torch.uint8, device="cuda")
    _lib.check(L.tt_actor_forward_store(ag.actor._h, obs.data_ptr(), 23, N, mu.data_ptr(), _lib.PRECISIONS[precision], C.byref(ring), s))
    _lib.check(L.tt_ou_step_store(x.data_ptr(), mu.data_ptr(), scaled.data_ptr(), N, 4, 0, ag.noise.iter_ptr, 0, C.byref(ring), s))
    _lib.check(L.tt_env_step_store(env._h, scaled.data_ptr(), obs2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), C.byref(ring), s))
    ref.store_transition(obs, mu, rew, obs2, done)
    for f in ("state_memory", "new_state_memory", "action_memory", "reward_memory", "terminal_memory"):
        assert torch.equal(getattr(fused, f), getattr(ref, f)), f


def test_cuda_graph_of_k_iterations_equals_eager_steps():
    """RolloutEngine.capture(): K iterations as ONE CUDA graph replay == the same K iterations launched one by one
    (trajectories, ring contents, statistics) -- the launch sequences of the C ABI are capturable as they are."""
    import ddpg_trucktrailer_b200 as tt
    N, cap = 2048, 8192

    def make():
        env = tt.VecTruckTrailerEnv(N, seed=5)
        ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=cap, actor_seed=1, precision="f16")
        eng = tt.RolloutEngine(env, ag)
        eng.reset(seed=5)
        return env, ag, eng

    e0, a0, g0 = make()
    e1, a1, g1 = make()
    k = g1.capture()                      # smallest K with K * N % cap == 0 and K even -> 4; runs 2 warm-up steps first
    assert k == 4
    for _ in range(2 + 3 * k):
        obs0, r0, d0 = g0.step()
    for _ in range(3):
        obs1, r1, d1 = g1.step_graph()
    assert g0.iterations == g1.iterations and a0.memory.mem_cntr == a1.memory.mem_cntr
    assert torch.equal(obs0, obs1) and torch.equal(r0, r1) and torch.equal(d0, d1)
    for name in ("state_memory", "action_memory", "reward_memory", "new_state_memory", "terminal_memory"):
        assert torch.equal(getattr(a0.memory, name), getattr(a1.memory, name)), name
    assert torch.equal(e0.get_state()["state"], e1.get_state()["state"])
    assert e0.read_stats() == e1.read_stats()
