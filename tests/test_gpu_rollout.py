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
