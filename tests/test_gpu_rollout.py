"""The whole rollout iteration (trainv2.py:511-531 without learn) through tt_rollout_step vs the same
sequence driven by hand through the reference-shaped Python API, and vs the oracle on a small batch."""
import os

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
        obs, _ = env2.reset(options={"mask": done})          # (auto_tick: the masked reset ends the iteration)
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


def test_step_out_buffers_receive_reward_and_done():
    """``RolloutEngine.step(out=(reward, done))``: the kernels write the step's results into the caller's buffers (two pairs
    used alternately by an asynchronous reader, bench.py e2e) -- same values as the engine's own buffers get."""
    import ddpg_trucktrailer_b200 as tt
    N, cap = 3000, 10000
    env1, ag1 = _mk(tt, N, cap); e1 = tt.RolloutEngine(env1, ag1); e1.reset()
    env2, ag2 = _mk(tt, N, cap); e2 = tt.RolloutEngine(env2, ag2); e2.reset()
    pairs = [(torch.full((N,), 7.0, device="cuda"), torch.full((N,), 9, dtype=torch.uint8, device="cuda")) for _ in range(2)]
    for it in range(40):
        o1, r1, d1 = e1.step()
        o2, r2, d2 = e2.step(out=pairs[it & 1])
        assert r2 is pairs[it & 1][0] and d2 is pairs[it & 1][1]
        assert torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(o1, o2), it
    for f in ("new_state_memory", "reward_memory", "terminal_memory"):
        assert torch.equal(getattr(ag1.memory, f), getattr(ag2.memory, f)), f
    with pytest.raises(ValueError):
        e2.step(out=(torch.zeros(N, device="cuda"), torch.zeros(N - 1, dtype=torch.uint8, device="cuda")))


@pytest.mark.parametrize("N", [77, 1000, 3000, 4096 + 33])
def test_done_bits_equal_packed_done_bytes(N):
    """``tt_env_set_done_bits``: every single-step launch also writes done as one bit per env (what a host that reads the flags
    back every step copies): equal to np.packbits(done, bitorder='little') for the rollout's kernel and for ``env.step``."""
    import ddpg_trucktrailer_b200 as tt
    env, ag = _mk(tt, N, 4 * N)
    eng = tt.RolloutEngine(env, ag); eng.reset()
    nw = (N + 31) // 32
    bits = [torch.full((nw,), -1, dtype=torch.int32, device="cuda") for _ in range(2)]
    rew = [torch.empty(N, device="cuda") for _ in range(2)]
    seen = 0
    for it in range(80):
        _, r, d = eng.step(out=(rew[it & 1], None, bits[it & 1]))
        want = np.packbits(d.cpu().numpy().astype(bool), bitorder="little")
        got = bits[it & 1].cpu().numpy().view(np.uint8)[:want.size]
        assert np.array_equal(got, want), it
        assert np.array_equal(tt.VecTruckTrailerEnv.unpack_done_bits(bits[it & 1], N), d.cpu().numpy().astype(bool))
        seen += int(d.sum())
    assert seen > 0
    # the plain step kernel (terminal observation, no reset) writes them too; None switches it off
    env.set_done_bits(bits[0])
    _, _, d, _ = env.step(torch.zeros(N, device="cuda"))
    assert np.array_equal(tt.VecTruckTrailerEnv.unpack_done_bits(bits[0], N), d.cpu().numpy())
    env.set_done_bits(None)
    bits[0].fill_(-1)
    env.reset(options={"mask": d}); env.step(torch.zeros(N, device="cuda"))
    assert int((bits[0] != -1).sum()) == 0


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


def test_flagship_rollout_vs_oracle_open_loop():
    """The BENCH path itself at BASELINE.json configs[1] scale: 65 536 envs, tcgen05 actor (f16), OU noise on, replay store
    fused into the producers, auto-reset, 120 iterations.  The raw actions are read back from the ring and replayed OPEN LOOP
    through the float64 oracle (adaptive RK45, reward_functionv1) from the same Philox poses, across auto-resets
    (north_star: identical initial states and action sequences).  Bars: observations of the whole trajectory (normalised
    states, sin / cos) 2e-5 absolute, final states 1e-4 relative, rewards 1e-4 relative, done exact except within-epsilon
    threshold crossings (every one checked against the oracle's margins), the ring's s / s' / r / done equal to what the
    oracle's loop would have stored, and the stored action = oracle actor(stored s) + oracle OU state within 1e-3."""
    from concurrent.futures import ThreadPoolExecutor
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    N, T, seed = 1 << 16, 120, 27
    cap = N * T
    env = tt.VecTruckTrailerEnv(N, seed=seed)
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=cap, actor_seed=0, seed=seed, precision="f16")
    sd = tt.init_actor_state_dict(seed=0); sd["mu.weight"] *= 20          # make the policy move the steering
    ag.load_actor_state_dict(sd)
    eng = tt.RolloutEngine(env, ag, store=True)
    eng.reset()
    for _ in range(T):
        eng.step()
    torch.cuda.synchronize()
    m = ag.memory
    assert m.mem_cntr == cap
    A = m.action_memory.view(T, N); R = m.reward_memory.view(T, N); D = m.terminal_memory.view(T, N)
    S = m.state_memory.view(T, N, 23); S2 = m.new_state_memory.view(T, N, 23)
    # ring consistency: s of iteration t + 1 is s' of iteration t unless the episode ended there (then it is a reset observation)
    keep = (D[:-1] == 0)
    assert torch.equal(S[1:][keep], S2[:-1][keep])
    final_state = env.state.cpu().numpy()
    chunk = 4096
    a_np, r_np, d_np = A.cpu().numpy(), R.cpu().numpy(), D.cpu().numpy()

    def run(i0):
        return orc.rollout_replay(seed, i0, a_np[:, i0:i0 + chunk])
    with ThreadPoolExecutor(os.cpu_count() or 4) as pool:
        outs = list(pool.map(run, range(0, N, chunk)))
    n_cross, n_done = 0, 0
    valids = []
    worst = dict(rew=0.0, obs=0.0, state=0.0)
    for ci, o in enumerate(outs):
        i0 = ci * chunk
        d_c = d_np[:, i0:i0 + chunk]
        # first step at which the two paths disagree about `done` (if any): must be a within-epsilon crossing; the env is
        # compared up to that step only (afterwards the two are in different episodes)
        valid = np.ones((T, chunk), bool)
        mis_t, mis_i = np.nonzero(d_c != o["done"])
        for i in np.unique(mis_i):
            t = mis_t[mis_i == i].min()
            # which flag differs is not stored in the ring: some raised oracle flag (or, if the oracle did not stop, any
            # threshold) must sit within its epsilon
            mg = o["margins"][t, i]
            near = [abs(mg[k]) < orc.MARGIN_EPS[k] for k in range(orc.NMARGINS)]
            assert any(near), f"env {i0 + i} step {t}: done {d_c[t, i]} vs oracle {o['done'][t, i]}, margins {mg}"
            valid[t:, i] = False
            n_cross += 1
        valids.append(valid)
        n_done += int(o["done"][valid].sum())
        rr = r_np[:, i0:i0 + chunk]
        worst["rew"] = max(worst["rew"], (np.abs(rr - o["rew"]) / np.maximum(np.abs(o["rew"]), 1.0))[valid].max())
        s_c = S[:, i0:i0 + chunk].cpu().numpy(); s2_c = S2[:, i0:i0 + chunk].cpu().numpy()
        worst["obs"] = max(worst["obs"], np.abs(s_c - o["s"])[valid].max(), np.abs(s2_c - o["s2"])[valid].max())
        last_ok = valid[-1] & (o["done"][-1] == 0)                       # (a finished env holds its NEXT episode's state)
        fs = final_state[i0:i0 + chunk][last_ok]; os_ = o["state"][-1][last_ok]
        worst["state"] = max(worst["state"], (np.abs(fs - os_) / np.maximum(np.abs(os_), 1.0)).max())
        if ci == 0:
            # the stored action: oracle actor on the stored observation + the oracle's OU state (512 envs x T rows)
            k = 512
            oa = orc.OracleActor({kk: v.cpu().numpy() for kk, v in ag.actor.state_dict().items()})
            mu = oa.forward(s_c[:, :k].reshape(-1, 23)).reshape(T, k)
            err = np.abs(a_np[:, :k] - (mu + o["ou"][:, :k]))[valid[:, :k]].max()
            assert err < 1e-3, err
    assert worst["rew"] < 1e-4 and worst["obs"] < 2e-5 and worst["state"] < 1e-4, worst
    assert n_cross <= N // 2000, n_cross
    assert n_done > N // 2                                               # the run crossed many auto-resets
    # all rows: stored action - OU state against a float64 torch forward of the stored observations
    sd64 = {kk: v.double() for kk, v in ag.actor.state_dict().items()}
    ou_all = torch.from_numpy(np.concatenate([o["ou"] for o in outs], 1)).cuda()
    ok_all = torch.from_numpy(np.concatenate(valids, 1)).cuda().reshape(-1)
    x = S.reshape(-1, 23).double()
    worst_mu = 0.0
    for j in range(0, x.shape[0], 1 << 20):
        h = torch.nn.functional.layer_norm(x[j:j + (1 << 20)] @ sd64["fc1.weight"].T + sd64["fc1.bias"], (400,), sd64["bn1.weight"], sd64["bn1.bias"]).relu()
        h = torch.nn.functional.layer_norm(h @ sd64["fc2.weight"].T + sd64["fc2.bias"], (300,), sd64["bn2.weight"], sd64["bn2.bias"]).relu()
        mu = torch.tanh(h @ sd64["mu.weight"].T + sd64["mu.bias"]).reshape(-1)
        got = (A.reshape(-1)[j:j + (1 << 20)] - ou_all.reshape(-1)[j:j + (1 << 20)]).double()
        worst_mu = max(worst_mu, float((got - mu).abs()[ok_all[j:j + (1 << 20)]].max()))
    assert worst_mu < 1e-3, worst_mu


@pytest.mark.parametrize("precision,N", [("fp32", 100), ("fp32", 3000), ("f16", 127), ("f16", 3000), ("f16", 148 * 128 * 2 + 77)])
@pytest.mark.parametrize("evaluate", [False, True])
def test_choose_action_one_launch_equals_separate_kernels(precision, N, evaluate):
    """tt_actor_choose_action (Agent.choose_action in one launch: the OU noise, the clip * pi / 4 scaling and the ring store
    of s and a live in the actor kernel's output stage) == tt_actor_forward_store + tt_ou_step_store, bit for bit: action,
    scaled action, OU state, ring rows."""
    import ctypes as C
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import _lib
    L = tt.load(); s = _lib.stream_ptr()
    cap, cntr0 = 1 << 16, 4100
    env = tt.VecTruckTrailerEnv(N, seed=9, global_env_offset=1000)
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=cap, actor_seed=2, precision=precision)
    sd = tt.init_actor_state_dict(seed=2); sd["mu.weight"] *= 50
    ag.load_actor_state_dict(sd)
    obs, _ = env.reset()
    env.tick(3)
    iter_ptr = L.tt_env_iter_ptr(env._h)
    prec = _lib.PRECISIONS[precision]
    x0 = torch.empty(N, device="cuda").normal_(0, 0.05)
    outs = []
    for fused in (False, True):
        mem = tt.DeviceReplayBuffer(cap)
        mem.state_memory.fill_(-7.0); mem.action_memory.fill_(-7.0)
        ring = _lib.ReplayRing(mem.state_memory.data_ptr(), mem.action_memory.data_ptr(), mem.reward_memory.data_ptr(),
                               mem.new_state_memory.data_ptr(), mem.terminal_memory.data_ptr(), cap, cntr0)
        a = torch.full((N,), -3.0, device="cuda"); sc = torch.full((N,), -3.0, device="cuda"); x = x0.clone()
        if fused:
            _lib.check(L.tt_actor_choose_action(ag.actor._h, obs.data_ptr(), 23, N, x.data_ptr(), 9, 1000, iter_ptr, int(evaluate),
                                                a.data_ptr(), sc.data_ptr(), prec, C.byref(ring), s))
        else:
            _lib.check(L.tt_actor_forward_store(ag.actor._h, obs.data_ptr(), 23, N, a.data_ptr(), prec, C.byref(ring), s))
            _lib.check(L.tt_ou_step_store(x.data_ptr(), a.data_ptr(), sc.data_ptr(), N, 9, 1000, iter_ptr, int(evaluate), C.byref(ring), s))
        outs.append((a, sc, x, mem.state_memory.clone(), mem.action_memory.clone()))
    for u, v, name in zip(outs[0], outs[1], ("action", "scaled", "ou_x", "ring.state", "ring.action")):
        assert torch.equal(u, v), name
    assert evaluate == bool(torch.equal(outs[1][2], x0))                 # the noise state moves exactly when noise is on


@pytest.mark.parametrize("N,cap,cntr0,store", [(3000, 10000, 9000, True), (5000, 2048, 77, True), (1 << 17, 1 << 18, 128, True), (4096, 4096, 0, False)])
def test_step_reset_one_launch_equals_separate_kernels(N, cap, cntr0, store):
    """tt_env_step_reset (env step + ring store + reset of finished envs + OU zeroing + iteration tick in ONE launch) ==
    tt_env_step_store + tt_env_reset(mask = done) + masked zeroing + tt_env_tick, bit for bit, over 40 iterations in which
    ~10 % of the envs finish an episode: observations (reset rows), rewards, done, ring rows (terminal rows), env state,
    start poses, OU state, statistics, iteration counter."""
    import ctypes as C
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import _lib
    L = tt.load(); s = _lib.stream_ptr()
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = torch.empty(40, N, device="cuda").uniform_(-0.785, 0.785, generator=g)
    res = []
    for fused in (False, True):
        env = tt.VecTruckTrailerEnv(N, seed=12, global_env_offset=77, auto_tick=False)
        obs, _ = env.reset()
        mem = tt.DeviceReplayBuffer(cap)
        for t_ in (mem.new_state_memory, mem.reward_memory): t_.fill_(-7.0)
        mem.terminal_memory.fill_(9)
        x = torch.ones(N, device="cuda")
        o2 = torch.zeros(N, 23, device="cuda"); rew = torch.zeros(N, device="cuda"); done = torch.zeros(N, dtype=torch.uint8, device="cuda")
        cntr, tot = cntr0, 0
        for t in range(40):
            ring = _lib.ReplayRing(mem.state_memory.data_ptr(), mem.action_memory.data_ptr(), mem.reward_memory.data_ptr(),
                                   mem.new_state_memory.data_ptr(), mem.terminal_memory.data_ptr(), cap, cntr)
            rp = C.byref(ring) if store else None
            if fused:
                _lib.check(L.tt_env_step_reset(env._h, acts[t].data_ptr(), o2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), x.data_ptr(), rp, s))
            else:
                if store:
                    _lib.check(L.tt_env_step_store(env._h, acts[t].data_ptr(), o2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), rp, s))
                else:
                    _lib.check(L.tt_env_step(env._h, acts[t].data_ptr(), o2.data_ptr(), 23, rew.data_ptr(), done.data_ptr(), None, s))
                _lib.check(L.tt_env_reset(env._h, done.data_ptr(), o2.data_ptr(), 23, s))
                x.masked_fill_(done.bool(), 0.0)
                _lib.check(L.tt_env_tick(env._h, 1, s))
            cntr += N; tot += int(done.sum())
            x += 0.25
        st = env.get_state()
        res.append(dict(obs=o2.clone(), rew=rew.clone(), done=done.clone(), s2=mem.new_state_memory.clone(), r=mem.reward_memory.clone(),
                        d=mem.terminal_memory.clone(), state=st["state"], start=st["start"], steps=st["episode_steps"], x=x.clone(),
                        stats=env.read_stats(), tot=tot))
    a, b = res
    assert a["tot"] == b["tot"] and a["tot"] > N // 20
    for k in ("obs", "rew", "done", "s2", "r", "d", "state", "start", "steps", "x"):
        assert torch.equal(a[k], b[k]), k
    for k in a["stats"]:      # counts are exact; the float sums depend on the (atomic) summation order
        if k in ("return_sum", "return_sq_sum", "reward_sum"):
            assert abs(a["stats"][k] - b["stats"][k]) <= 1e-9 * abs(b["stats"][k]) + 1e-6, k
        else:
            assert a["stats"][k] == b["stats"][k], k


def test_auto_precision_picks_the_kernel_by_batch_size():
    """TT_PREC_AUTO (north_star (c)): warp-level fp32 FMA below the crossover, tcgen05 tiles from it on.  Which kernel ran is
    observable: both are deterministic, so the auto output is bit-identical to the explicit mode's output."""
    import ddpg_trucktrailer_b200 as tt
    actor = tt.agent.CudaActor(); actor.load_state_dict(tt.init_actor_state_dict(seed=4))
    for n, want in ((1, "fp32"), (64, "fp32"), (191, "fp32"), (192, "f16"), (4096, "f16"), (1 << 18, "f16")):
        assert actor.auto_precision(n) == want
        obs = torch.empty(n, 23, device="cuda").uniform_(-1, 1)
        assert torch.equal(actor.forward(obs, precision="auto").clone(), actor.forward(obs, precision=want).clone())
    other = tt.agent.CudaActor(23, 256, 128); other.load_state_dict(tt.init_actor_state_dict(23, 256, 128, 1, seed=3))
    assert other.auto_precision(1 << 20) == "fp32"                       # the tensor-core kernel is specialised to 23-400-300
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=8, max_size=64)
    assert ag.precision == "auto"
    with pytest.raises(ValueError):
        tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=8, max_size=64, precision="bf16")
    with pytest.raises(ValueError):
        actor.forward(torch.zeros(4, 23, device="cuda"), precision="f16_plain")
    with pytest.raises(ValueError):
        ag.choose_action(torch.zeros(9, 23, device="cuda"))               # one observation row per environment


def test_full_size_rollout_properties():
    """BASELINE.json's full size (2^22 envs per GPU), the bench path itself (tcgen05 actor + OU noise + fused ring store + in-kernel
    reset, two launches per iteration): size-independent properties instead of an oracle replay.  (1) what the ring holds is what
    the kernels handed over: new_state == the next observation for every env that continues, state(t + 1) == the observation the
    actor of t + 1 read, action / reward / terminal rows == the step's outputs; (2) clip * pi / 4 is exact; (3) the bit-packed flags
    equal the bytes; (4) the device statistics count every step; (5) a 4 096-env shard with the matching global offset reproduces
    its slice of rewards and flags bit for bit (Philox streams are keyed by the global env id, the actor is position independent)."""
    import ddpg_trucktrailer_b200 as tt
    N, cap, K, warm = 1 << 22, 1 << 23, 4, 40
    sd = tt.init_actor_state_dict(seed=0); sd["mu.weight"] *= 100

    def mk(n, offset, ring):
        env = tt.VecTruckTrailerEnv(n, seed=13, global_env_offset=offset)
        ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=n, max_size=ring, actor_seed=0, seed=13, global_env_offset=offset, precision="f16")
        ag.load_actor_state_dict(sd)
        eng = tt.RolloutEngine(env, ag); eng.reset()
        return env, ag, eng

    env, ag, eng = mk(N, 0, cap)
    S = 4096
    env_s, ag_s, eng_s = mk(S, N - S, 4 * S)
    for _ in range(warm):
        eng.step(); eng_s.step()
    env.read_stats()
    m = ag.memory
    bits = torch.zeros((N + 31) // 32, dtype=torch.int32, device="cuda")
    env.set_done_bits(bits)
    prev_obs, ndone = None, 0
    for t in range(K):
        c0 = m.mem_cntr % m.mem_size
        obs, r, d = eng.step()
        _, r_s, d_s = eng_s.step()
        rows = slice(c0, c0 + N)                                        # cap is a multiple of N: no wrap inside an iteration
        db = d.bool()
        assert torch.equal(m.reward_memory[rows], r) and torch.equal(m.terminal_memory[rows].bool(), db)
        assert torch.equal(m.action_memory[rows].reshape(-1), eng.action)
        assert torch.equal(eng.scaled, eng.action.clamp(-1, 1) * np.float32(0.78539819))
        ns = m.new_state_memory[rows]
        assert torch.equal(ns[~db], obs[~db]) and not torch.equal(ns[db], obs[db])          # finished envs: terminal row vs reset observation
        if prev_obs is not None:
            assert torch.equal(m.state_memory[rows], prev_obs)
        prev_obs = obs.clone()
        got = bits.cpu().numpy().view(np.uint8)
        assert np.array_equal(got, np.packbits(db.cpu().numpy(), bitorder="little"))
        assert torch.equal(r[N - S:], r_s) and torch.equal(d[N - S:], d_s)
        assert torch.isfinite(r).all()
        ndone += int(db.sum())
    env.set_done_bits(None)
    st = env.read_stats()
    assert st["steps"] == N * K and st["episodes"] == ndone and ndone > N // 200
