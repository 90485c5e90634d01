"""Row f1: the hand-written CUDA learner (tt_learn_step, csrc/tt_learn.cu) against the reference's Agent.learn
(DDPG/DDPG_agent.py:72-131): golden parameters written by the untouched reference (tests/golden/ref_learn.npz), per-tensor
gradients against torch autograd in float64, sampling, graph capture, and the hand-over of the new policy to the rollout
actor."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
NETS = ("actor", "target_actor", "critic", "target_critic")


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_learn.npz"))
    sd = lambda phase, net: {k.split("/", 2)[2]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{phase}/{net}/")}
    return g, sd


def _agent_from_golden(tt, g, sd, cap=None):
    dims = tuple(int(v) for v in g["dims"])
    alpha, beta, tau, gamma = (float(v) for v in g["hyper"])
    B, steps = int(g["batch"]), int(g["steps"])
    ag = tt.VecAgent(alpha, beta, (dims[0],), tau, 1, gamma=gamma, max_size=cap or B * steps, fc1_dims=dims[1], fc2_dims=dims[2],
                     batch_size=B, num_envs=B, precision="fp32")
    ag.load_actor_state_dict(sd("before", "actor"))
    t = lambda k: torch.from_numpy(g[k]).cuda()
    for j in range(steps):                                    # the reference's transitions, in the reference's order
        sl = slice(j * B, (j + 1) * B)
        ag.remember(t("s")[sl], t("a")[sl].reshape(-1), t("r")[sl], t("s2")[sl], torch.from_numpy(g["d"][sl].astype(np.uint8)).cuda())
    ln = ag.learner
    for n in NETS:
        ln.load_state_dict(n, sd("before", n))
    return ag, ln, B, steps


def test_learn_step_matches_reference_learn(golden_dir):
    """Three consecutive Agent.learn() calls of the UNTOUCHED reference (fc1 = 64, fc2 = 48, batch 64; the reference's initial
    weights and batches) vs three tt_learn_step calls on the same ring rows: all four networks within rtol 1e-5 (atol 2e-7:
    Adam divides by sqrt(v) + 1e-8, so a gradient that is pure rounding noise moves its parameter by up to lr = 1e-3 / 1e-4 in
    either implementation; there are none in this fixture)."""
    import ddpg_trucktrailer_b200 as tt
    g, sd = _golden(golden_dir)
    ag, ln, B, steps = _agent_from_golden(tt, g, sd)
    for j in range(steps):
        ln.learn(rows=torch.arange(j * B, (j + 1) * B))
    worst = {}
    for n in NETS:
        want, got = sd("after", n), ln.state_dict(n)
        for k in want:
            w = want[k].cuda()
            err = (got[k] - w).abs()
            worst[f"{n}.{k}"] = float((err / (1e-5 * w.abs() + 2e-7)).max())
        moved = max((sd("before", n)[k] - want[k]).abs().max().item() for k in want)
        assert moved > (1e-7 if n.startswith("target") else 1e-5)
    bad = {k: v for k, v in worst.items() if v > 1.0}
    assert not bad, bad
    # the new policy was handed to the rollout actor (re-packed inside tt_learn_step)
    obs = torch.empty(300, 23, device="cuda").uniform_(-1, 1)
    ref = tt.agent.CudaActor(*ag.actor.dims); ref.load_state_dict(sd("after", "actor"))
    assert (ag.actor.forward(obs) - ref.forward(obs)).abs().max() < 1e-5
    assert all(torch.allclose(ag.actor.state_dict()[k], sd("after", "actor")[k].cuda(), rtol=1e-5, atol=2e-7) for k in tt.ACTOR_KEYS)


@pytest.mark.parametrize("dims,B", [((23, 400, 300), 64), ((23, 64, 48), 64), ((23, 130, 77), 37)])
def test_learner_gradients_match_autograd_float64(dims, B):
    """Every gradient tensor of one update (critic loss w.r.t. the critic, actor loss w.r.t. the actor through the UPDATED
    critic) against torch autograd in float64 on the same batch; then the parameters after the step against the torch
    restatement (oracle/torch_learner.py, pinned on the reference) in float32."""
    import ddpg_trucktrailer_b200 as tt
    from oracle.torch_learner import TorchLearner, _Actor, _Critic
    torch.manual_seed(3)
    cap = 512
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=cap, fc1_dims=dims[1], fc2_dims=dims[2], batch_size=B, num_envs=cap,
                     precision="fp32", actor_seed=11)
    s = torch.empty(cap, 23, device="cuda").uniform_(-1, 1); s2 = torch.empty(cap, 23, device="cuda").uniform_(-1, 1)
    a = torch.empty(cap, device="cuda").uniform_(-1.2, 1.2); r = torch.empty(cap, device="cuda").normal_(5, 30)
    d = (torch.rand(cap, device="cuda") < 0.15).to(torch.uint8)
    ag.remember(s, a, r, s2, d)
    ln = ag.learner
    # make LayerNorm affine parameters and the heads non-trivial
    for n in ("actor", "critic"):
        sd_ = ln.state_dict(n)
        for k in sd_:
            if k.startswith("bn"):
                sd_[k] = sd_[k] + torch.empty_like(sd_[k]).uniform_(-0.3, 0.3)
        for k in ("mu.weight", "q.weight"):
            if k in sd_:
                sd_[k] = sd_[k] * 30
        ln.load_state_dict(n, sd_); ln.load_state_dict("target_" + n, {k: v * 0.97 for k, v in sd_.items()})
    before = {n: ln.state_dict(n) for n in NETS}
    rows = torch.randint(0, cap, (B,), device="cuda")
    # float32 torch restatement on the same rows
    class FakeMem:
        mem_cntr = cap
        def sample_buffer(self, bs): return s[rows], a[rows].reshape(-1, 1), r[rows], s2[rows], d[rows].bool()
    dims_t = tuple(dims)

    class FakeActor:
        dims = dims_t
        def state_dict(self): return before["actor"]
        def load_state_dict(self, sd_): pass
    class FakeAgent:
        device = torch.device("cuda")
        actor, memory = FakeActor(), FakeMem()
        alpha, beta, tau, gamma, batch_size = 1e-4, 1e-3, 1e-3, 0.99, B
    tl = TorchLearner(FakeAgent())
    for n in NETS:
        getattr(tl, n).load_state_dict(before[n])
    # float64 autograd gradients
    A64, TA64, C64, TC64 = (_Actor(*dims).double().cuda(), _Actor(*dims).double().cuda(), _Critic(*dims).double().cuda(), _Critic(*dims).double().cuda())
    for net, n in ((A64, "actor"), (TA64, "target_actor"), (C64, "critic"), (TC64, "target_critic")):
        net.load_state_dict({k: v.double() for k, v in before[n].items()})
    sb, ab, rb, s2b, db = s[rows].double(), a[rows].double().reshape(-1, 1), r[rows].double(), s2[rows].double(), d[rows].bool()
    with torch.no_grad():
        q2 = TC64(s2b, TA64(s2b)).masked_fill(db.view(-1, 1), 0.0)
        y = (rb + 0.99 * q2.view(-1)).view(B, 1)
    torch.nn.functional.mse_loss(y, C64(sb, ab)).backward()
    gc64 = {k: p.grad.clone() for k, p in C64.named_parameters()}
    ln.learn(rows=rows)
    tl.learn()
    gc = ln.grads("critic")
    for k in gc64:
        scale = gc64[k].abs().max().item() + 1e-30
        assert (gc[k].double() - gc64[k]).abs().max().item() < 2e-5 * scale, ("critic", k)
    # actor gradient through the UPDATED critic (float64 copy of the CUDA learner's new critic)
    C64.load_state_dict({k: v.double() for k, v in ln.state_dict("critic").items()})
    (-C64(sb, A64(sb))).mean().backward()
    ga = ln.grads("actor")
    for k, p in A64.named_parameters():
        scale = p.grad.abs().max().item() + 1e-30
        assert (ga[k].double() - p.grad).abs().max().item() < 2e-5 * scale, ("actor", k)
    # parameters after the step vs the float32 torch restatement.  Adam normalises by sqrt(v): after ONE step the update is
    # lr * g / (|g| + 1e-8), i.e. +-lr wherever |g| >> 1e-8 regardless of rounding in g
    for n in NETS:
        want, got = getattr(tl, n).state_dict(), ln.state_dict(n)
        for k in want:
            assert torch.allclose(got[k], want[k], rtol=2e-5, atol=2e-6), (n, k, (got[k] - want[k]).abs().max().item())
    assert (ln.state_dict("critic")["fc1.weight"] - before["critic"]["fc1.weight"]).abs().max() > 5e-4


def test_learner_samples_uniformly_is_deterministic_and_graph_capturable():
    """replay_buffer.py:23-34: rows uniform with replacement over the FILLED part of the ring (Philox stream keyed by the
    update counter: a new batch every step, the same sequence for the same seed); a CUDA graph of the launch sequence
    replays to the same parameters as eager calls."""
    import ctypes as C
    import ddpg_trucktrailer_b200 as tt
    cap, fill, B = 4096, 1000, 64

    def make():
        ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, max_size=cap, batch_size=B, num_envs=fill, precision="fp32", actor_seed=2, seed=5)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        u = lambda *sh: torch.empty(*sh, device="cuda").uniform_(-1, 1, generator=g)
        ag.remember(u(fill, 23), u(fill), u(fill) * 20, u(fill, 23), (u(fill) > 0.8).to(torch.uint8))
        torch.manual_seed(9)
        return ag, ag.learner

    ag1, l1 = make()
    seen = []
    for _ in range(40):
        l1.learn()
        off = l1.L.tt_learner_last_rows(l1._h) - l1._ws.data_ptr()
        seen.append(l1._ws[off:off + 8 * B].view(torch.int64).clone())
    rows = torch.stack(seen)
    assert rows.min() >= 0 and rows.max() < fill                         # only the filled part
    assert not torch.equal(rows[0], rows[1])                             # a fresh batch every step
    u = rows.double() / fill
    assert abs(u.mean().item() - 0.5) < 0.02 and abs(u.var().item() - 1 / 12) < 0.01
    # same seed, same ring -> same trajectory of parameters; here as ONE captured graph replayed 40 times
    a3, b3 = make()
    b3.learn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        b3.learn()
    # the capture recorded one update without running it: 39 replays = updates 2..40
    for _ in range(39):
        gr.replay()
    torch.cuda.synchronize()
    for n in NETS:
        x, y = l1.state_dict(n), b3.state_dict(n)
        for k in x:
            assert torch.equal(x[k], y[k]), (n, k)
    obs = torch.empty(256, 23, device="cuda").uniform_(-1, 1)
    assert torch.equal(ag1.actor.forward(obs).clone(), a3.actor.forward(obs).clone())      # the re-pack ran inside the graph too


def test_async_trainer_hides_the_update_and_keeps_semantics():
    """rollout.AsyncTrainer: one rollout iteration + one DDPG update per step with the update on a side stream.  Checks the
    contract: (1) every update samples only ring rows that earlier iterations completed (never the rows being written), also
    after the ring has wrapped; (2) the rollout of iteration t + 1 runs with the policy the update of iteration t produced
    (the packed actor equals the learner's parameters of one step ago); (3) the learner made one update per iteration."""
    import ddpg_trucktrailer_b200 as tt
    N, cap, iters = 4096, 4096 * 5, 14
    env = tt.VecTruckTrailerEnv(N, seed=3)
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=cap, actor_seed=1, seed=3, precision="f16")
    eng = tt.RolloutEngine(env, ag, store=True)
    eng.reset()
    tr = tt.AsyncTrainer(eng, reserve_sms=2)
    ln = tr.ln
    seen_flat = []
    ref = tt.agent.CudaActor()
    probe = torch.empty(300, 23, device="cuda").uniform_(-1, 1)
    for it in range(iters):
        c0 = ag.memory.mem_cntr
        begin, count = tr._window()
        tr.step()
        torch.cuda.synchronize()
        if it > 0 and count >= 64:
            off = ln.L.tt_learner_last_rows(ln._h) - ln._ws.data_ptr()
            rows = ln._ws[off:off + 8 * 64].view(torch.int64).cpu().numpy()
            written = (c0 + np.arange(N)) % cap                         # the rows iteration `it` was writing
            assert not np.intersect1d(rows, written).size, it
            rel = (rows - begin) % cap
            assert rel.max() < count, it
            if c0 < cap:
                assert rows.max() < c0                                   # before the wrap: only rows that exist
        seen_flat.append(ln._flat["actor"].clone())
        if it >= 2:
            # iteration `it` ran with the actor packed from the learner's parameters after update it - 1 (the packed images are
            # what counts: state_dict() of a load_flat actor is a live view of the learner's vector)
            ref.load_flat(seen_flat[it - 1])
            for prec in ("fp32", "f16"):
                assert torch.equal(ag.actor.forward(probe, precision=prec).clone(), ref.forward(probe, precision=prec).clone()), (it, prec)
    assert tr.updates == iters - 1
    assert not torch.equal(seen_flat[1], seen_flat[-1])                  # the policy moved
    tr.close()
