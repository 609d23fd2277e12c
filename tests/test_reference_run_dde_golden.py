"""The oracle's delay path against golden vectors produced by the REFERENCE'S OWN, UNMODIFIED code
(tools/make_reference_dde_golden.py: `interpolation/interpolate*.py`, `xde/base_dde.py`, `functional/ddeint.py` and the
fixed solvers executed on the NumPy `paddle` stand-in of oracle/ref_shim -- this part of HEAD runs without repairs).

Bit for bit: `evaluate` / `derivative` of the three interpolants (uniform and non-uniform grids, queries inside, on
grid points and outside the span), `HistoryIndex.forward`, `HistoryIndex.backward` (the gradient of the lags), and whole
`ddeint` solves with Euler / Midpoint / RK4 through `BaseDDE.move` / the damped `fuse`."""
import os

import numpy as np
import pytest

from tests.problems import dde_field_coefficients

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_dde_vectors.npz"), allow_pickle=False)
GATHER = sorted({k.split("/")[1] for k in Z.files if k.startswith("gather/")})
DDEINT = sorted({k.split("/")[1] for k in Z.files if k.startswith("ddeint/") and k.count("/") == 2})
f32 = np.float32


@pytest.mark.parametrize("kind", ["linear", "cubic", "bez"])
@pytest.mark.parametrize("name", GATHER)
def test_oracle_gather_reproduces_the_reference_interpolants(oracle, name, kind):
    v, d = oracle.history_gather(kind, Z[f"gather/{name}/his"], Z[f"gather/{name}/span"], Z[f"gather/{name}/lags"])
    assert np.array_equal(v, Z[f"gather/{name}/{kind}/val"]), "evaluate"
    assert np.array_equal(d, Z[f"gather/{name}/{kind}/der"]), "derivative"


def test_oracle_reproduces_history_index_forward_and_backward(oracle):
    v, d = oracle.history_gather("cubic", Z["index/his"], Z["index/span"], Z["index/lags"])
    assert np.array_equal(v, Z["index/y_lags"])
    assert np.array_equal(oracle.history_gather_bwd(Z["index/grad_y"], d), Z["index/grad_lags"])


def oracle_ddeint(oracle, method, t, y0, y_lags):
    """FixedSolver.integrate (base_fixed_solver.py:103-144) on grid == t_span, stepping through the oracle's damped fuse;
    fp32 time arithmetic as the reference's tensors."""
    ca, cb = dde_field_coefficients()
    move = lambda y: y_lags * ca - y * cb  # noqa: E731
    fuse = lambda dy, dt, y: oracle.dde_fuse(dy, float(dt), y)  # noqa: E731
    third = f32(1 / 3)
    y, sol = y0, [y0]
    for i in range(1, t.size):
        dt = f32(t[i] - t[i - 1])
        if method == "euler":                       # fixed_solver/euler.py:7-11
            y = fuse(move(y), dt, y)
        elif method == "midpoint":                  # fixed_solver/midpoint.py:7-18
            half = f32(0.5) * dt
            y = fuse(move(fuse(move(y), half, y)), dt, y)
        else:                                       # rk4_alt_step_func, base_fixed_solver.py:166-197
            k1 = move(y)
            k2 = move(fuse(k1, dt * third, y))
            k3 = move(fuse(k1 - k2 * third, dt, y))
            k4 = move(fuse(k1 - k2 + k3, dt, y))
            y = (fuse(k1, dt, y) + f32(3) * fuse(k2, dt, y) + f32(3) * fuse(k3, dt, y) + fuse(k4, dt, y)) * f32(0.125)
        sol.append(y)
    return np.concatenate(sol, axis=-2)


@pytest.mark.parametrize("name", DDEINT)
def test_oracle_reproduces_the_reference_ddeint(oracle, name):
    y_lags, _ = oracle.history_gather("cubic", Z["index/his"], Z["index/span"], Z["index/lags"])
    out = oracle_ddeint(oracle, name.split("_")[0], Z[f"ddeint/{name}/t"], Z["ddeint/y0"], y_lags)
    ref = Z[f"ddeint/{name}/sol"]
    assert out.shape == ref.shape and np.array_equal(out, ref), float(np.abs(out - ref).max())


def test_committed_dde_vectors_are_what_the_reference_computes_here():
    from oracle.ref_shim import loader

    if not loader.available():
        pytest.skip("/root/reference is not on this machine")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_reference_dde_golden", os.path.join(ROOT, "tools", "make_reference_dde_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out, _ = gen.generate()
    assert sorted(out) == sorted(Z.files)
    for k in Z.files:
        a, b = np.asarray(out[k]), Z[k]
        assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), k
