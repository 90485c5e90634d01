"""The algebra the tensor-core actor kernel relies on (csrc/tt_actor_tc4.cu: pack kernels and epilogues), restated in float64
numpy and checked against the oracle's ActorNetwork.forward (DDPG/networks.py:138-147) -- no GPU needed:
  * LayerNorm 1 folded into W1 (centring + scale) with the variance from the Cholesky factor of the centred Gram matrix,
  * LayerNorm 2's centring folded into W2, the variance as a plain sum of squares,
  * relu(y) = (y + |y|) / 2 with the y / 2 half of the output dot as one linear GEMM column,
  * w3 relu(g z + be) = |z + be / g| (w3 |g| / 2) + linear part, incl. g < 0 and g == 0 columns."""
import os

import numpy as np


def _folded_forward(w, obs):
    f = lambda k: w[k].astype(np.float64)
    W1, b1, g1, be1 = f("fc1.weight"), f("fc1.bias"), f("bn1.weight"), f("bn1.bias")
    W2, b2, g2, be2 = f("fc2.weight"), f("fc2.bias"), f("bn2.weight"), f("bn2.bias")
    w3, b3 = f("mu.weight").ravel(), float(w["mu.bias"][0])
    X = np.concatenate([obs.astype(np.float64), np.ones((len(obs), 1))], 1)           # constant-1 column carries the bias
    # layer 1: image rows = g1 (Wf - m); statistic rows = Cholesky factor of the centred Gram matrix
    Wf = np.concatenate([W1, b1[:, None]], 1)
    Wc = Wf - Wf.mean(0)
    G = Wc.T @ Wc
    L = np.linalg.cholesky(G + 1e-18 * np.eye(len(G)))
    var1 = ((X @ L) ** 2).sum(1) / W1.shape[0]                                       # |L^T x|^2 / 400 = var(h)
    a2 = np.maximum((X @ (g1[:, None] * Wc).T) / np.sqrt(var1 + 1e-5)[:, None] + be1, 0.0)
    # layer 2: centred weights, linear column, |.| epilogue
    A = np.concatenate([a2, np.ones((len(obs), 1))], 1)
    Wf2 = np.concatenate([W2, b2[:, None]], 1)
    W2c = Wf2 - Wf2.mean(0)
    h = A @ W2c.T                                                                    # = h2 - mean(h2)
    lin = A @ (0.5 * (W2c * (g2 * w3)[:, None]).sum(0))
    rstd = 1.0 / np.sqrt((h ** 2).mean(1) + 1e-5)
    ok = np.abs(g2) > 1e-30
    E = np.where(ok, be2 / np.where(ok, g2, 1.0), 0.0)
    Wh = np.where(ok, 0.5 * w3 * np.abs(g2), 0.0)
    c0 = b3 + np.where(ok, 0.5 * be2, np.maximum(be2, 0.0)) @ w3
    return np.tanh(c0 + rstd * lin + (np.abs(h * rstd[:, None] + E) * Wh).sum(1))


def test_folded_actor_equals_oracle_forward(golden_dir):
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "ref_actor.npz"))
    w = {k[3:]: g[k].copy() for k in g.files if k.startswith("w0/")}
    rng = np.random.default_rng(11)
    w["mu.weight"] = w["mu.weight"] * np.float32(40.0)
    w["bn1.weight"] = rng.uniform(0.5, 1.5, 400).astype(np.float32)
    w["bn1.bias"] = rng.uniform(-0.2, 0.2, 400).astype(np.float32)
    gw = rng.uniform(0.5, 1.5, 300).astype(np.float32)
    gw[rng.choice(300, 40, replace=False)] *= -1
    gw[rng.choice(300, 20, replace=False)] = 0.0
    w["bn2.weight"] = gw
    w["bn2.bias"] = rng.uniform(-0.3, 0.3, 300).astype(np.float32)
    obs = rng.uniform(-2, 2, (777, 23)).astype(np.float32)
    ref = orc.OracleActor(w).forward(obs)
    assert np.abs(_folded_forward(w, obs) - ref).max() < 2e-6                        # the oracle computes in float32
    assert np.abs(_folded_forward(w, g["obs"]) - orc.OracleActor(w).forward(g["obs"])).max() < 2e-6
