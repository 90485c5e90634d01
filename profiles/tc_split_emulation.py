"""CPU emulation (float64 arithmetic, operands rounded to fp16 exactly where the tensor-core kernel rounds them) of the actor
forward with different first-layer operand splits, on the amplified "trained-like" weight set of tests/golden/ref_actor.npz:
does a cheaper first layer (2 products instead of 3) still hold north_star's 1e-3 bar?
    python profiles/tc_split_emulation.py > profiles/r02_tc_split_emulation.txt"""
import os, sys
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
g = np.load(os.path.join(root, "tests", "golden", "ref_actor.npz"))
h16 = lambda a: a.astype(np.float16).astype(np.float64)


def forward(w, obs, mode):
    X = np.concatenate([obs.astype(np.float64), np.ones((len(obs), 1))], 1)              # constant-1 column carries fc1.bias
    Wf = np.concatenate([w["fc1.weight"], w["fc1.bias"][:, None]], 1).astype(np.float64)
    m = Wf.mean(0)
    W1 = w["bn1.weight"].astype(np.float64)[:, None] * (Wf - m)                           # LayerNorm 1 folded into the operand
    Xh, Wh = h16(X), h16(W1)
    Xl, Wl = h16(X - Xh), h16(W1 - Wh)
    t = {"3 products (TT_PREC_F16)": Xh @ Wh.T + Xl @ Wh.T + Xh @ Wl.T,
         "2 products, X exact": Xh @ Wh.T + Xl @ Wh.T,
         "2 products, W exact": Xh @ Wh.T + Xh @ Wl.T,
         "1 product (TT_PREC_F16_PLAIN)": Xh @ Wh.T,
         "exact first layer": X @ W1.T}[mode]
    h = X @ Wf.T                                                                          # statistics: exact (Cholesky columns are split too)
    rstd = 1.0 / np.sqrt(h.var(1) + 1e-5)
    a2 = h16(np.maximum(t * rstd[:, None] + w["bn1.bias"], 0.0))
    W2 = np.concatenate([w["fc2.weight"], w["fc2.bias"][:, None]], 1).astype(np.float64)
    W2c = h16(W2 - W2.mean(0))                                                            # centred over the outputs, fp16 image
    A2 = np.concatenate([a2, np.ones((len(obs), 1))], 1)
    y = A2 @ W2c.T
    y = y / np.sqrt((y * y).mean(1, keepdims=True) + 1e-5) * w["bn2.weight"] + w["bn2.bias"]
    return np.tanh(np.maximum(y, 0.0) @ w["mu.weight"][0].astype(np.float64) + w["mu.bias"][0])


w0 = {k[3:]: g[k] for k in g.files if k.startswith("w0/")}
w1 = {k: v.copy() for k, v in w0.items()}                                                # the amplified set of tests/test_gpu_agent.py::_sets
w1["mu.weight"] = w1["mu.weight"] * np.float32(60.0); w1["mu.bias"] = w1["mu.bias"] + np.float32(0.05)
w1["fc2.weight"] = w1["fc2.weight"] * np.float32(2.0)
for k in ("bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"):
    w1[k] = g["w1/" + k]
for name, w, ref in (("reference init", w0, g["out0"]), ("amplified", w1, g["out1"])):
    for mode in ("exact first layer", "3 products (TT_PREC_F16)", "2 products, X exact", "2 products, W exact", "1 product (TT_PREC_F16_PLAIN)"):
        err = np.abs(forward(w, g["obs"], mode) - ref)
        print(f"{name:15s} {mode:32s} max |err| = {err.max():.3e}   rms = {np.sqrt((err ** 2).mean()):.3e}")
