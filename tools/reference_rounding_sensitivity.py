"""How much do the reference's results depend on HOW AN EAGER OP ROUNDS -- the one thing the NumPy `paddle` stand-in
decides and Paddle does not document?

The golden vectors (tests/golden/reference_run_*.npz) are produced with every op rounding as the arithmetic
specification says.  This tool re-runs the same reference code on the same inputs with a DIFFERENT,
equally plausible op-level arithmetic -- `x ** (1/p)` through float32 `pow`, `mean()` / `sum()` through NumPy's pairwise
fp32 reductions, and the caller's field through BLAS matmuls + libm `tanh` with its vector-Jacobian product in plain
NumPy -- and reports, per case, whether the accept / reject sequence is the same and how far step sizes, solutions and
gradients move.  It bounds what "parity against Paddle itself" could look like for any implementation that is not
Paddle's own binary: the north star's rtol 1e-5 on results, and an exact step sequence only where the decisions are
not within rounding noise of the threshold.

    python tools/reference_rounding_sensitivity.py          (needs /root/reference; prints a table)"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import loader  # noqa: E402

f32 = np.float32


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class NumpyField:
    """tanh(pre(y) @ W1 + b1) @ W2 + b2 with BLAS matmuls and libm tanh; VJP in plain NumPy (fp32)."""

    def __init__(self, w1, b1, w2, b2, pre):
        self.w1, self.b1, self.w2, self.b2, self.pre = w1, b1, w2, b2, pre
        self.d, self.h = w1.shape

    def _pre(self, y):
        return {"id": y, "square": y * y, "cube": y * y * y}[self.pre]

    def _dpre(self, y):
        return {"id": np.ones_like(y), "square": f32(2) * y, "cube": f32(3) * y * y}[self.pre]

    def __call__(self, t, y):
        y = np.asarray(y, f32).reshape(-1, self.d)
        return (np.tanh(self._pre(y) @ self.w1 + self.b1) @ self.w2 + self.b2).astype(f32)

    def vjp_batch(self, y, c):
        y, c = np.asarray(y, f32).reshape(-1, self.d), np.asarray(c, f32).reshape(-1, self.d)
        u = self._pre(y)
        h = np.tanh(u @ self.w1 + self.b1)
        f = h @ self.w2 + self.b2
        dz = (c @ self.w2.T) * (f32(1) - h * h)
        dy = (dz @ self.w1.T) * self._dpre(y)
        return f.astype(f32), dy.astype(f32), [(u.T @ dz).astype(f32), dz.sum(0).astype(f32), (h.T @ c).astype(f32),
                                               c.sum(0).astype(f32)]


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def compare_logs(lg, ref):
    same_len = len(lg) == len(ref)
    same_seq = same_len and np.array_equal(lg["accepted"], ref["accepted"])
    ddt = float(np.abs(lg["dt"] / ref["dt"] - 1).max()) if same_seq else float("nan")
    return same_seq, len(ref), len(lg), ddt


def main():
    assert loader.available(), "needs the reference tree"
    fwd, adj = _tool("make_reference_golden"), _tool("make_reference_adjoint_golden")
    zf = np.load(fwd.OUT, allow_pickle=False)
    za = np.load(adj.OUT, allow_pickle=False)
    os.environ["XDE_SHIM_ROUNDING"] = "libm"
    ns = loader.load()
    assert ns.paddle.ROUNDING == "libm"
    rows = []
    for name, kind, solver, d, h, pre, B, t, opts in fwd.cases():
        if kind != "adaptive":
            continue
        w = [zf[f"{name}/{k}"] for k in ("w1", "b1", "w2", "b2")]
        sol, log = fwd.run_reference(ns, kind, solver, NumpyField(*w, pre), zf[f"{name}/y0"], t, opts)
        same, n_ref, n_alt, ddt = compare_logs(log, zf[f"{name}/log"])
        rows.append((f"fwd {name}", opts.get("rtol", 1e-7), same, n_ref, n_alt, ddt, rel(sol, zf[f"{name}/sol"]), None))
    for name, d, h, pre, B, t, opts, adj_norm, t_grad in adj.cases():
        w = [za[f"{name}/{k}"] for k in ("w1", "b1", "w2", "b2")]
        sol, gp, a0, gt, log = adj.run_reference(ns, NumpyField(*w, pre), w, za[f"{name}/y0"], t, opts, adj_norm, t_grad,
                                                 za[f"{name}/grad_y"])
        same, n_ref, n_alt, ddt = compare_logs(log, za[f"{name}/log"])
        g = np.concatenate([x.ravel() for x in gp])
        g_ref = np.concatenate([za[f"{name}/{k}"].ravel() for k in ("gw1", "gb1", "gw2", "gb2")])
        rows.append((f"adj {name} ({adj_norm})", opts.get("rtol", 1e-7), same, n_ref, n_alt, ddt,
                     max(rel(sol, za[f"{name}/sol"]), rel(a0, za[f"{name}/adj_y0"])), rel(g, g_ref)))
    print(f"{'case':46s} {'rtol':>7s} {'same accept/reject seq':>22s} {'attempts spec/alt':>18s} {'max|dt/dt-1|':>13s} "
          f"{'states rel':>11s} {'grads rel':>10s}")
    for name, rtol, same, n_ref, n_alt, ddt, rs, rg in rows:
        print(f"{name:46s} {rtol:7.0e} {str(same):>22s} {n_ref:>9d}/{n_alt:<8d} {ddt:13.2e} {rs:11.2e} "
              f"{'' if rg is None else format(rg, '10.2e')}")
    n_same = sum(r[2] for r in rows)
    print(f"\n{n_same} of {len(rows)} cases keep the accept/reject sequence; largest state difference "
          f"{max(r[6] for r in rows):.2e}, largest gradient difference {max(r[7] for r in rows if r[7] is not None):.2e} "
          f"(relative to the largest entry)")


if __name__ == "__main__":
    main()
