"""The oracle against golden vectors produced by the REFERENCE'S OWN, UNMODIFIED solver code (tools/
make_reference_golden.py: /root/reference's files executed on the NumPy `paddle` stand-in of oracle/ref_shim).

Bit for bit: every solution, and for the adaptive solvers the whole attempt log (t0, dt, error ratio, accepted) --
all five tableaux, solver options, min_step, step_t / jump_t, B = 1 and B > 1 (the reference's global norm = the
oracle's controller="batch"), the fixed solvers with both interpolants and step_size / grid_constructor grids.
This pins the oracle's restatement of the reference's control flow and formulas to what the reference's code computes
("adjoint, SDE and step sequences are necessarily parity unpinned" -- step sequences no longer are).  The
eager ops themselves follow the repository's arithmetic specification in the stand-in: see its docstring."""
import ast
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_run_vectors.npz")
RK = {"Dopri5": "dopri5", "Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "AdaptiveHeun": "adaptive_heun", "Dopri8": "dopri8"}


def load_cases():
    z = np.load(GOLD, allow_pickle=False)
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


Z, NAMES = load_cases()


def oracle_solution(oracle, z, name):
    meta = ast.literal_eval(str(z[f"{name}/meta"]))
    om = oracle.MLP(z[f"{name}/w1"], z[f"{name}/b1"], z[f"{name}/w2"], z[f"{name}/b2"], pre=meta["pre"])
    y0, t = z[f"{name}/y0"], z[f"{name}/t"]
    if meta["kind"] == "adaptive":
        kw = {k: v for k, v in meta.items() if k not in ("kind", "solver", "pre")}
        out, st, log, rc = oracle.adaptive_rk_mlp(RK[meta["solver"]], om, y0, t, controller="batch", **kw)
        assert rc == 0
        return out, log, meta
    grid = t
    if "step_size" in meta:  # _grid_constructor_from_step_size (base_fixed_solver.py:66-89) in fp32
        step = np.float32(meta["step_size"])
        niters = int(np.ceil(np.float32((t[-1] - t[0]) / step) + np.float32(1.0)))
        grid = np.arange(0, niters, dtype=np.float32) * step + t[0]
        grid[-1] = t[-1]
    if "grid" in meta:
        grid = np.asarray(meta["grid"], np.float32)
    g = np.ascontiguousarray(grid[:t.size])
    y = oracle.fixed_mlp(meta["solver"].lower(), om, y0, g)  # [B, T, D] on the grid the solver really steps on
    out = np.empty_like(y)
    out[:, 0] = y[:, 0]
    for i in range(1, t.size):  # linear_interp (interp_fn.py:4-10); cubic on grid == t_span is the identity too
        if t[i] == g[i - 1]:
            out[:, i] = y[:, i - 1]
        elif t[i] == g[i]:
            out[:, i] = y[:, i]
        else:
            slope = np.float32(np.float32(t[i] - g[i - 1]) / np.float32(g[i] - g[i - 1]))
            out[:, i] = y[:, i - 1] + slope * (y[:, i] - y[:, i - 1])
    return out, None, meta


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_reference_run(oracle, name):
    out, log, meta = oracle_solution(oracle, Z, name)
    ref = Z[f"{name}/sol"]
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), f"{name}: max|d| = {np.abs(out - ref).max()}"
    if meta["kind"] == "adaptive":
        rlog = Z[f"{name}/log"]
        assert len(log) == len(rlog)
        for f in ("t0", "dt", "ratio", "accepted"):
            assert np.array_equal(log[f], rlog[f]), f"{name}: attempt log field {f}"


def test_committed_vectors_are_what_the_reference_computes_here():
    """When the reference tree is present (the build container; never the GPU box) re-run its code and compare with the
    committed file: the fixtures cannot drift from the reference."""
    from oracle.ref_shim import loader

    if not loader.available():
        pytest.skip("/root/reference is not on this machine")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(ROOT, "tools", "make_reference_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out, _ = gen.generate()
    assert sorted(out) == sorted(Z.files)
    for k in Z.files:
        a, b = np.asarray(out[k]), Z[k]
        assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), k
