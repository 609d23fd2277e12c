"""Diagnostic: large-state adjoint gradient error per parameter block, several shapes / T / B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200._lib as L
if len(sys.argv) > 1:
    L._SO = os.path.abspath(sys.argv[1])
import paddlexde_b200 as px
from oracle import xde_oracle as xo
from paddlexde_b200.functional.odeint_adjoint import adjoint_backward
from tests.problems import fanin_weights

def run(d, h, B, T, pre="id", seed=1, gy_mode="last"):
    w = fanin_weights(d, h, seed=seed)
    field, om = px.MLPField(*w, pre=pre), xo.MLP(*w, pre=pre)
    y0 = np.random.default_rng(3).uniform(-1, 1, (B, d)).astype(np.float32)
    t = np.linspace(0, 1, T).astype(np.float32)
    kw = dict(rtol=1e-6, atol=1e-8)
    ref, _, _, rc = xo.dopri5_mlp(om, y0, t, **kw)
    gy = np.zeros_like(ref)
    gy[-1] = np.sign(ref[-1]) / ref[-1].size
    if gy_mode == "all":
        gy = (np.random.default_rng(5).standard_normal(ref.shape) / ref[0].size).astype(np.float32)
    g, a0, stats, _ = adjoint_backward(field, t, ref, gy, return_adj_y0=True, **kw)
    g_ref, a_ref, st_ref, _, rc = xo.dopri5_mlp_adjoint(om, t, ref, gy, **kw)
    g = px._tensor.to_host(g)
    s = stats.read()
    parts = om.split(g), om.split(g_ref)
    errs = [float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)) for a, b in zip(*parts)]
    rej = int((st_ref.n_attempts - st_ref.n_accepted).sum())
    print(f"d={d} h={h} B={B} T={T} {pre} gy={gy_mode}: a0 exact={np.array_equal(px._tensor.to_host(a0), a_ref)} attempts "
          f"{s.n_attempts}/{int(st_ref.n_attempts.sum())} rejections={rej}  rel err gW1,gb1,gW2,gb2 = "
          + " ".join(f"{e:.2e}" for e in errs), flush=True)

for args in [(64, 256, 32, 2), (64, 256, 1, 2), (64, 256, 32, 3), (64, 256, 70, 4), (32, 64, 64, 2), (32, 64, 129, 4), (16, 64, 200, 4),
             (64, 128, 33, 4), (32, 128, 130, 4), (32, 256, 65, 4)]:
    run(*args)
run(64, 256, 70, 4, gy_mode="all")
run(32, 64, 64, 2, pre="cube")
