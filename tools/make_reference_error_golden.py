"""The reference's ASSERTIONS on the adaptive path, produced by its own code (unmodified solver files on the NumPy
`paddle` stand-in): which inputs make `AdaptiveRKSolver.step` / `_adaptive_step` raise, with which message
(solver/base_adaptive_solver_rk.py:120-122, 200-203), after how many attempts, and what the solution rows completed
before the assertion hold.

    python tools/make_reference_error_golden.py   -> tests/golden/reference_run_errors.npz (+ a summary)

tests/test_reference_run_error_golden.py: the oracle reports the matching status word on the same inputs, after the same
number of attempts, with the same completed rows; the package's shim turns that status word back into an
`AssertionError` with the reference's message."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xde_oracle as xo  # noqa: E402
from oracle.ref_shim import loader  # noqa: E402
from tests.problems import cfg2_tspan, cfg2_y0, fanin_weights, spiral_weights  # noqa: E402

f32 = np.float32
OUT = os.path.join(ROOT, "tests", "golden", "reference_run_errors.npz")


def cases():
    """(name, weights, pre, y0, t_span, solver options)"""
    big = [f32(40.0) * a for a in fanin_weights(2, 16, seed=3)]
    bad = cfg2_y0(5, seed=2)
    bad[3, 1] = np.inf
    return [("max_num_steps", spiral_weights(), "cube", cfg2_y0(7, seed=1), np.array([0.0, 2.0, 5.0], f32), dict(max_num_steps=3)),
            ("max_num_steps_b1", spiral_weights(), "cube", cfg2_y0(1, seed=3), np.array([0.0, 5.0], f32), dict(max_num_steps=2)),
            ("nonfinite_initial_state", spiral_weights(), "cube", bad, cfg2_tspan(4), {}),
            # y**3 overflows fp32: the field returns NaN from a finite state
            ("field_returns_nan", spiral_weights(), "cube", np.array([[1e13, -1e13]], f32), cfg2_tspan(3), {}),
            ("field_returns_nan_first_step_given", big, "cube", np.array([[1e13, -1e13], [0.5, 0.25]], f32), cfg2_tspan(3),
             dict(first_step=0.01)),
            ("dt_underflow", spiral_weights(), "cube", cfg2_y0(1, seed=4), np.array([1e8, 1e8 + 64.0], f32), dict(first_step=1.0))]


def run_reference(ns, om, y0, t, opts):
    P = ns.paddle
    func = lambda tt, y: P.Tensor(om(0.0, np.ascontiguousarray(y.a, f32)).reshape(y.a.shape))  # noqa: E731
    tT, y0T = P.to_tensor(t, dtype=P.float32), P.to_tensor(y0, dtype=P.float32)
    n = [0]
    try:
        s = ns.Dopri5(xde=ns.BaseODE(func, y0=y0T, t_span=tT), y0=y0T, rtol=opts.get("rtol", 1e-7), atol=opts.get("atol", 1e-9),
                      norm=ns.ode_utils._rms_norm, **{k: v for k, v in opts.items() if k not in ("rtol", "atol")})
        s.func = func
        orig = s._adaptive_step

        def counted(state):
            n[0] += 1
            return orig(state)

        s._adaptive_step = counted
        # AdaptiveSolver.integrate (base_adaptive_solver.py:24-31), row by row so that the rows completed before the
        # assertion can be kept
        rows = [np.array(y0, f32)]
        s._before_integrate(tT)
        for i in range(1, len(t)):
            rows.append(np.ascontiguousarray(s.step(tT[i]).a, f32))
        return None, n[0], np.stack(rows)
    except AssertionError as e:
        return str(e), n[0], np.stack(rows) if "rows" in dir() else np.array(y0, f32)[None]


def generate():
    ns = loader.load()
    out, summary = {}, []
    for name, w, pre, y0, t, opts in cases():
        w = [np.ascontiguousarray(a, f32) for a in w]
        msg, n_calls, rows = run_reference(ns, xo.MLP(*w, pre=pre), y0, t, opts)
        assert msg is not None, f"{name}: the reference did not raise"
        out[f"{name}/w1"], out[f"{name}/b1"], out[f"{name}/w2"], out[f"{name}/b2"] = w
        out[f"{name}/y0"], out[f"{name}/t"], out[f"{name}/rows_done"] = y0, t, rows
        out[f"{name}/message"] = np.array(msg[:40])
        out[f"{name}/adaptive_step_calls"] = np.array(n_calls)
        out[f"{name}/meta"] = np.array(repr(dict(pre=pre, **opts)))
        summary.append(f"{name:26s} AssertionError({msg[:44]!r}...) after {n_calls} _adaptive_step calls, {len(rows)} rows done")
    return out, summary


if __name__ == "__main__":
    xo.build()
    out, summary = generate()
    np.savez_compressed(OUT, **out)
    print("\n".join(summary))
    print("wrote", OUT)
