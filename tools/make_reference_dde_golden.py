"""Golden vectors of the delay path produced by the REFERENCE'S OWN CODE on the NumPy `paddle` stand-in of oracle/ref_shim:
  * `interpolation/interpolate_base.py` + `interpolate.py` (LinearInterpolation, CubicHermiteSpline, BezierSpline:
    `evaluate` and `derivative`) on uniform and non-uniform grids, queries inside, on grid points and outside the span;
  * `xde/base_dde.py`: `HistoryIndex.forward` / `.backward` (the gradient with respect to the lags) and `BaseDDE.fuse`;
  * `functional/ddeint.py` + `solver/fixed_solver/{euler,midpoint,rk4}.py`: whole `ddeint` solves.
All of these files are unmodified and nothing had to be repaired: this part of HEAD runs as it is.

    python tools/make_reference_dde_golden.py   -> tests/golden/reference_run_dde_vectors.npz (+ a summary)

The stand-in's op-level rounding (what is NOT the reference's): `@` = rounded products summed left to right in fp32,
`x ** 3 = (x * x) * x`, `sum(axis=[0, 1, 3])` accumulated in fp64 -- see oracle/ref_shim/paddle/__init__.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import loader  # noqa: E402
from tests.problems import dde_field_coefficients  # noqa: E402

f32 = np.float32
OUT = os.path.join(ROOT, "tests", "golden", "reference_run_dde_vectors.npz")
KINDS = {"linear": "LinearInterpolation", "cubic": "CubicHermiteSpline", "bez": "BezierSpline"}


def gather_cases():
    """(name, his shape, grid, queries)"""
    rng = np.random.default_rng(11)
    uni = np.arange(24, dtype=f32)
    non = np.cumsum(rng.uniform(0.2, 1.5, 40)).astype(f32)
    q_uni = np.concatenate([(np.arange(12) + rng.uniform(0, 1, 12)), [0.0, 1.0, 22.0, 22.5, 23.0, 3.25, -0.75, 24.5]]).astype(f32)
    q_non = np.concatenate([rng.uniform(non[0], non[-1], 21), [non[0], non[7], non[-2], non[-1], non[0] - 0.4, non[-1] + 2.0]]).astype(f32)
    return [("uniform_4d", (2, 5, 24, 3), uni, q_uni),       # cfg5's layout [batch, nodes, T, D]
            ("nonuniform_3d", (4, 40, 7), non, q_non),
            ("short_grid", (3, 4, 2), np.array([0.0, 0.5, 2.0, 2.25], f32), np.array([0.1, 0.5, 1.9, 2.25, 3.0, -1.0], f32)),
            ("one_row_wide", (1, 33, 64), np.linspace(0, 8, 33).astype(f32), rng.uniform(0, 8, 9).astype(f32))]


def generate():
    ns = loader.load()
    P = ns.paddle
    out, summary = {}, []
    rng = np.random.default_rng(12)
    for name, shape, grid, q in gather_cases():
        his = rng.standard_normal(shape).astype(f32)
        out[f"gather/{name}/his"], out[f"gather/{name}/span"], out[f"gather/{name}/lags"] = his, grid, q
        for kind, cls in KINDS.items():
            it = getattr(ns.interpolate, cls)(P.to_tensor(his), P.to_tensor(grid))
            out[f"gather/{name}/{kind}/val"] = np.ascontiguousarray(it.evaluate(P.to_tensor(q)).a, f32)
            out[f"gather/{name}/{kind}/der"] = np.ascontiguousarray(it.derivative(P.to_tensor(q)).a, f32)
        summary.append(f"gather/{name}: his {shape}, {q.size} queries, 3 interpolants")
    # HistoryIndex (cubic, the PyLayer's default) forward + backward: D3STN's layout, learnable real-valued lags
    his = rng.uniform(-1, 1, (3, 7, 48, 3)).astype(f32)
    span = np.arange(48, dtype=f32)
    lags = (np.arange(12) * 3 + rng.uniform(0, 1, 12)).astype(f32)
    y_lags = ns.base_dde.HistoryIndex.apply(lags=P.to_tensor(lags), his=P.to_tensor(his), his_span=P.to_tensor(span))
    gy = rng.standard_normal(y_lags.a.shape).astype(f32)
    g_lags, g_his, g_span = ns.base_dde.HistoryIndex.backward(y_lags._ctx, P.to_tensor(gy))
    assert g_his is None and g_span is None
    out["index/his"], out["index/span"], out["index/lags"], out["index/grad_y"] = his, span, lags, gy
    out["index/y_lags"], out["index/grad_lags"] = np.ascontiguousarray(y_lags.a, f32), np.ascontiguousarray(g_lags.a, f32)
    summary.append(f"index: HistoryIndex forward {y_lags.a.shape}, backward -> grad_lags {g_lags.a.shape}")
    # ddeint: the reference's entry point with its own Euler / Midpoint / RK4 stepping through BaseDDE.move / fuse
    ca, cb = dde_field_coefficients()
    func = lambda yl, y: yl * float(ca) - y * float(cb)  # noqa: E731  (Tensor ops of the stand-in: fp32 elementwise)
    y0 = rng.uniform(-1, 1, (3, 7, 12, 3)).astype(f32)
    for solver, t, interp in (("Euler", np.arange(2, dtype=f32), ""), ("Euler", np.array([0, 0.5, 1.25, 2.0], f32), "linear"),
                              ("Midpoint", np.array([0, 0.5, 1.25], f32), "linear"), ("RK4", np.array([0.1, 0.4, 1.0], f32), "cubic")):
        sol, yl = ns.ddeint(func, P.to_tensor(y0), P.to_tensor(t), P.to_tensor(lags), P.to_tensor(his), P.to_tensor(span),
                            getattr(ns, solver), fixed_solver_interp=interp)
        key = f"ddeint/{solver.lower()}_T{t.size}"
        out[f"{key}/t"], out[f"{key}/sol"] = t, np.ascontiguousarray(sol.a, f32)
        out[f"{key}/interp"] = np.array(interp)
        assert np.array_equal(yl.a, out["index/y_lags"])
        summary.append(f"{key}: solution {sol.a.shape} (interp={interp!r})")
    out["ddeint/y0"] = y0
    return out, summary


if __name__ == "__main__":
    out, summary = generate()
    np.savez_compressed(OUT, **out)
    print("\n".join(summary))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
