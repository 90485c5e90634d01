"""Actor forward / OU noise / action scaling kernels vs the reference torch outputs (golden) and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sets(g):
    w0 = {k[3:]: g[k] for k in g.files if k.startswith("w0/")}
    w1 = {k: v.copy() for k, v in w0.items()}
    w1["mu.weight"] = w1["mu.weight"] * np.float32(60.0)
    w1["mu.bias"] = w1["mu.bias"] + np.float32(0.05)
    w1["fc2.weight"] = w1["fc2.weight"] * np.float32(2.0)
    for k in ("bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"):
        w1[k] = g["w1/" + k]
    return w0, w1


@pytest.mark.parametrize("ld", [23, 24])
def test_actor_fp32_matches_reference_torch(golden_dir, ld):
    """ActorNetwork.forward (networks.py:138-147): fp32 kernel within 1e-5 of the reference torch forward."""
    import ddpg_trucktrailer_b200 as tt
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    obs = torch.zeros(len(g["obs"]), ld, device="cuda")
    obs[:, :23] = torch.from_numpy(g["obs"]).cuda()
    actor = tt.agent.CudaActor()
    for w, ref in zip(_sets(g), (g["out0"], g["out1"])):
        actor.load_state_dict(w)
        out = actor.forward(obs[:, :23]).cpu().numpy()
        assert np.abs(out - ref).max() < 1e-5


@pytest.mark.parametrize("n", [1, 7, 8, 9, 63, 64, 65, 191, 256, 257, 1000, 20000])
def test_actor_fp32_ragged_sizes_vs_oracle(golden_dir, n):
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    _, w1 = _sets(g)
    obs = np.random.default_rng(n).uniform(-1, 1, (n, 23)).astype(np.float32)
    actor = tt.agent.CudaActor(); actor.load_state_dict(w1)
    out = actor.forward(torch.from_numpy(obs).cuda()).cpu().numpy()
    assert np.abs(out - orc.OracleActor(w1).forward(obs)).max() < 1e-5


def test_actor_other_hidden_sizes():
    """fc1_dims / fc2_dims are constructor arguments of the reference Agent (DDPG_agent.py:10-12)."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    sd = tt.init_actor_state_dict(23, 256, 128, 1, seed=3)
    sd["mu.weight"] *= 50
    actor = tt.agent.CudaActor(23, 256, 128); actor.load_state_dict(sd)
    obs = np.random.default_rng(0).uniform(-1, 1, (777, 23)).astype(np.float32)
    out = actor.forward(torch.from_numpy(obs).cuda()).cpu().numpy()
    ref = orc.OracleActor({k: v.numpy() for k, v in sd.items()}).forward(obs)
    assert np.abs(out - ref).max() < 1e-5


@pytest.mark.parametrize("dims", [(23, 256, 128), (23, 130, 77), (23, 512, 512), (23, 33, 31)])
@pytest.mark.parametrize("n", [1, 5, 70])
def test_actor_fp32_small_batches_other_hidden_sizes(dims, n):
    """Small batches take the cluster-split fp32 kernel (one eighth of the hidden units per CTA, statistics and activations
    exchanged through distributed shared memory): padded and ragged unit slices, slices without any real unit."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    sd = tt.init_actor_state_dict(*dims, 1, seed=4)
    sd["mu.weight"] *= 30
    actor = tt.agent.CudaActor(*dims); actor.load_state_dict(sd)
    obs = np.random.default_rng(n).uniform(-1, 1, (n, 23)).astype(np.float32)
    out = actor.forward(torch.from_numpy(obs).cuda(), precision="fp32").cpu().numpy()
    ref = orc.OracleActor({k: v.numpy() for k, v in sd.items()}).forward(obs)
    assert np.abs(out - ref).max() < 1e-5


def test_init_matches_reference_distributions():
    import ddpg_trucktrailer_b200 as tt
    sd = tt.init_actor_state_dict(seed=0)
    assert sd["fc1.weight"].abs().max() <= 1 / np.sqrt(400) and sd["fc2.weight"].abs().max() <= 1 / np.sqrt(300)
    assert sd["mu.weight"].abs().max() <= 0.003 and (sd["bn1.weight"] == 1).all() and (sd["bn2.bias"] == 0).all()


def test_init_equals_reference_seed0(golden_dir):
    """Same draw order as networks.py:110-131 -> torch.manual_seed(0) reproduces the reference's weights."""
    import ddpg_trucktrailer_b200 as tt
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    sd = tt.init_actor_state_dict(seed=0)
    for k in tt.ACTOR_KEYS:
        assert np.array_equal(sd[k].numpy(), g["w0/" + k]), k


def test_ou_noise_vs_oracle_and_moments():
    """noise.py:12-17 with Philox normals: matches the oracle restatement; stationary moments are right."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    from ddpg_trucktrailer_b200 import _lib
    L = _lib.load()
    n = 1 << 16
    x = torch.zeros(n, device="cuda"); act = torch.full((n,), 0.25, device="cuda")
    it = torch.tensor([7], dtype=torch.int32, device="cuda")
    xo = np.zeros(n, np.float32); ao = np.full(n, 0.25, np.float32)
    for t in (7, 8, 9):
        it.fill_(t)
        _lib.check(L.tt_ou_step(x.data_ptr(), act.data_ptr(), None, n, 99, 1000, it.data_ptr(), _lib.stream_ptr()))
        orc.ou_step(xo, ao, None, 99, 1000, t)
    assert np.abs(x.cpu().numpy() - xo).max() < 2e-6 and np.abs(act.cpu().numpy() - ao).max() < 5e-6
    # reset mask zeroes the state first (agent.noise.reset(), trainv2.py:492)
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda"); mask[::2] = 1
    it.fill_(10)
    _lib.check(L.tt_ou_step(x.data_ptr(), None, mask.data_ptr(), n, 99, 1000, it.data_ptr(), _lib.stream_ptr()))
    m = mask.cpu().numpy().astype(np.uint8)
    orc.ou_step(xo, None, m, 99, 1000, 10)
    assert np.abs(x.cpu().numpy() - xo).max() < 2e-6
    inc = x.cpu().numpy()[::2]                      # one step from 0: N(0, (0.15*0.1)^2)
    assert abs(inc.mean()) < 3e-4 and abs(inc.std() - 0.015) < 3e-4


def test_choose_action_and_scaling():
    """Agent.choose_action (DDPG_agent.py:36-49): mu + noise unclipped; evaluate=True skips noise;
    trainv2.py:516 scaling."""
    import ddpg_trucktrailer_b200 as tt
    from ddpg_trucktrailer_b200 import _lib
    N = 5000
    ag = tt.VecAgent(1e-4, 1e-3, (23,), 1e-3, 1, num_envs=N, max_size=1 << 14, actor_seed=0)
    obs = torch.empty(N, 23, device="cuda").uniform_(-1, 1)
    mu = ag.choose_action(obs, evaluate=True).clone()
    assert mu.shape == (N, 1)
    a1 = ag.choose_action(obs).clone()
    assert torch.allclose(a1 - mu, ag.noise.x_prev.reshape(-1, 1), atol=1e-7) and (a1 != mu).any()
    scaled = torch.empty(N, device="cuda")
    big = (a1 * 100).reshape(-1).contiguous()
    _lib.check(_lib.load().tt_scale_action(big.data_ptr(), scaled.data_ptr(), N, _lib.stream_ptr()))
    ref = np.clip(big.cpu().numpy(), -1, 1) * np.float32(0.78539819)
    assert np.array_equal(scaled.cpu().numpy(), ref.astype(np.float32))
    ag.noise.reset()
    assert (ag.noise.x_prev == 0).all()


# ------------------------------------------------------------------------------------------------------------
# tcgen05 tensor-core actor: north_star bar 1e-3 against the reference torch forward
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol0,tol1", [("f16", 1e-4, 1e-3), ("f16_plain", 1e-4, 2e-3), ("bf16", 1e-3, 2e-2)])
def test_actor_tc_matches_reference_torch(golden_dir, precision, tol0, tol1):
    """Set 0 = the reference's own initialisation (torch.manual_seed(0)); set 1 = deliberately amplified
    "trained-like" weights (mu.weight x60, fc2 x2, random LayerNorm affine).  The default tensor-core mode
    ("f16": fp16 operands, exact split first layer) holds 1e-3 on both; plain fp16 ("f16_plain", 10 % faster) and plain
    bf16 operands hold it on the reference-scale weights only (their error on set 1 is pure operand rounding)."""
    import ddpg_trucktrailer_b200 as tt
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    obs = torch.from_numpy(g["obs"]).cuda()
    actor = tt.agent.CudaActor()
    for w, ref, tol in zip(_sets(g), (g["out0"], g["out1"]), (tol0, tol1)):
        actor.load_state_dict(w)
        out = actor.forward(obs, precision=precision, allow_out_of_bar=True).cpu().numpy()
        err = np.abs(out - ref).max()
        assert err < tol, (precision, err)


@pytest.mark.parametrize("precision", ["f16", "f16_plain", "bf16"])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 148 * 128 * 3 + 17])
def test_actor_tc_ragged_sizes_vs_fp32_kernel(golden_dir, n, precision):
    import ddpg_trucktrailer_b200 as tt
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    w0, _ = _sets(g)
    actor = tt.agent.CudaActor(); actor.load_state_dict(w0)
    obs = torch.empty(n, 24, device="cuda").uniform_(-1, 1)[:, :23]          # ld_obs = 24
    a = actor.forward(obs, precision=precision, allow_out_of_bar=True).clone()
    b = actor.forward(obs, precision="fp32")
    assert (a - b).abs().max() < (1e-3 if precision == "bf16" else 1e-4)
    a2 = actor.forward(obs, precision=precision, allow_out_of_bar=True)
    assert torch.equal(a, a2)                                                # deterministic


def test_actor_tc_refuses_other_layer_sizes():
    """The tensor-core kernel is specialised to 23-400-300; other sizes must fail loudly, not fall back."""
    import ddpg_trucktrailer_b200 as tt
    actor = tt.agent.CudaActor(23, 256, 128); actor.load_state_dict(tt.init_actor_state_dict(23, 256, 128, 1, seed=3))
    with pytest.raises(tt.TTError):
        actor.forward(torch.zeros(4, 23, device="cuda"), precision="f16")


@pytest.mark.parametrize("precision", ["f16", "bf16"])
def test_actor_tc_full_size_position_independence(golden_dir, precision):
    """BASELINE size (2^22 rows): a row's output must not depend on where the row sits -- tile, TMEM lane, CTA, CTA-pair rank,
    ring slot phase.  The batch is a 4 099-row block (not a multiple of the 128-row tile) repeated to 2^22 rows: every copy of
    the block must be bit-identical to the first one, and the first one within the bar of the fp32 kernel."""
    import ddpg_trucktrailer_b200 as tt
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    _, w1 = _sets(g)
    actor = tt.agent.CudaActor(); actor.load_state_dict(w1)
    N, B = 1 << 22, 4099
    block = torch.empty(B, 23, device="cuda").uniform_(-2, 2)
    reps = (N + B - 1) // B
    obs = block.repeat(reps, 1)[:N].contiguous()
    out = actor.forward(obs, precision=precision, allow_out_of_bar=True)
    ref = actor.forward(block, precision="fp32")
    assert (out[:B] - ref).abs().max() < (2e-2 if precision == "bf16" else 1.5e-3)
    full = out[: (N // B) * B].view(N // B, B)
    assert torch.equal(full, full[0:1].expand_as(full))
    assert torch.equal(out[(N // B) * B:], out[: N - (N // B) * B])
    # CTA pairs with the multicast W2 stream (batches above 2^18 rows) against single CTAs (the block alone): same bits
    assert torch.equal(out[:B], actor.forward(block, precision=precision, allow_out_of_bar=True))


@pytest.mark.parametrize("precision,tol", [("f16", 1e-3), ("f16_plain", 1.5e-3)])
def test_actor_tc_zero_and_negative_layernorm2_weights(golden_dir, precision, tol):
    """The tensor-core epilogue rewrites w3 relu(g z + be) as |z + be / g| (w3 |g| / 2) plus a linear GEMM column: columns with
    g == 0 (constant relu(be) w3), g < 0 and tiny |g| must still match the oracle."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    _, w1 = _sets(g)
    w1 = {k: v.copy() for k, v in w1.items()}
    rng = np.random.default_rng(5)
    gw = w1["bn2.weight"]
    gw[rng.choice(300, 40, replace=False)] *= -1
    gw[rng.choice(300, 25, replace=False)] = 0.0
    gw[rng.choice(300, 10, replace=False)] = 1e-35
    gw[7] = 3e-6
    w1["bn2.bias"] = rng.uniform(-0.3, 0.3, 300).astype(np.float32)
    obs = rng.uniform(-1, 1, (4099, 23)).astype(np.float32)
    actor = tt.agent.CudaActor(); actor.load_state_dict(w1)
    ref = orc.OracleActor(w1).forward(obs)
    assert np.abs(actor.forward(torch.from_numpy(obs).cuda()).cpu().numpy() - ref).max() < 1e-5
    assert np.abs(actor.forward(torch.from_numpy(obs).cuda(), precision=precision, allow_out_of_bar=True).cpu().numpy() - ref).max() < tol


def test_actor_tc_rank_deficient_first_layer(golden_dir):
    """LayerNorm 1's variance comes out of the layer-1 GEMM as |L^T x|^2 with L the Cholesky factor of the centred Gram matrix of
    [fc1.weight | fc1.bias] (computed by one warp at tt_actor_load).  Inputs the first layer ignores (zero columns), duplicated
    columns and a constant bias make that matrix singular: the zero-pivot columns of L must be dropped, not divided by."""
    import ddpg_trucktrailer_b200 as tt
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    w0, _ = _sets(g)
    w = {k: v.copy() for k, v in w0.items()}
    w["fc1.weight"][:, 3] = 0.0                       # an input the network ignores
    w["fc1.weight"][:, 7] = w["fc1.weight"][:, 5]     # two identical columns
    w["fc1.weight"][:, 11] = -2.0 * w["fc1.weight"][:, 9]
    w["fc1.bias"][:] = 0.25                           # constant bias: its centred column is zero
    rng = np.random.default_rng(11)
    obs = rng.uniform(-2, 2, (1000, 23)).astype(np.float32)
    actor = tt.agent.CudaActor(); actor.load_state_dict(w)
    ref = orc.OracleActor(w).forward(obs)
    x = torch.from_numpy(obs).cuda()
    assert np.abs(actor.forward(x, precision="fp32").cpu().numpy() - ref).max() < 1e-5
    out = actor.forward(x, precision="f16").cpu().numpy()
    assert np.isfinite(out).all() and np.abs(out - ref).max() < 1e-3
