"""Generates the committed fixtures under tests/golden/ (run from the repo root: python tools/make_golden.py).

Two kinds, kept in separate files because they pin different things:

* reference_fixtures.npz -- the known answers the REFERENCE's own tests hold for this path, evaluated from the
  closed forms in /root/reference/tests/testing_utils.py:8-70 (Sine / Constant / Linear problems at
  t = linspace(1, 8, 10)) and tests/interpolation/test_interpolation.py:13-85 (ramp / sin interpolation
  targets).  They pin the ORACLE (tests/test_golden.py, CPU) and, through it, the kernels.
* oracle_vectors.npz -- outputs of the C oracle on seeded inputs for every family of the path (the reference
  itself cannot run here: no Paddle; "parity unpinned").  They freeze the arithmetic
  specification: a later edit of oracle/ or of a kernel that changes one bit of these fails the suite, and
  the GPU tests compare the kernels with them WITHOUT calling the oracle.

Inputs are regenerated from the seeds in tests/problems.py; only outputs (and small inputs) are stored."""
import math
import os
import sys

import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xde_oracle as xo  # noqa: E402
from tests.problems import cfg2_tspan, cfg2_y0, fanin_weights, spiral_weights  # noqa: E402

f32 = np.float32
OUT = os.path.join(ROOT, "tests", "golden")


def reference_fixtures():
    t = np.linspace(1, 8, 10).astype(f32)                     # construct_problem, testing_utils.py:86
    t64 = t.astype(np.float64)
    sine = (-0.5 * t64 ** 4 * np.cos(2 * t64) + 0.5 * t64 ** 3 * np.sin(2 * t64) + 0.25 * t64 ** 2 * np.cos(2 * t64)
            - t64 ** 3 + 2 * t64 ** 4 + (math.pi - 0.25) * t64 ** 2)      # SineXDE.y_exact :33-41
    const = 0.2 * t64 + 3.0                                                # ConstantXDE.y_exact :25-26
    rng = np.random.default_rng(0)
    U = rng.standard_normal((10, 10)) * 0.1
    A = (2 * U - (U + U.T)).astype(f32)                                     # LinearXDE :48-49 (seeded here)
    lin = np.stack([scipy.linalg.expm(A.astype(np.float64) * float(ti)) @ np.ones(10) for ti in t])
    return dict(t=t, sine=sine, constant=const, linear_A=A, linear=lin,
                interp_ramp_t=f32(21.12), interp_ramp_val=np.array([21.12 * 0.5, 0.0]), interp_ramp_der=np.array([0.5, 0.0]),
                interp_sin_t=f32(16.5), interp_sin_val=np.array([math.sin(16.5), 0.0]),
                interp_sin_der=np.array([math.cos(16.5), 0.0]))


def oracle_vectors():
    g = {}
    om = xo.MLP(*spiral_weights(), pre="cube")
    # cfg1 / cfg2 family: dopri5 forward + adjoint, B = 64 of the cfg2 batch
    y0, t = cfg2_y0(64), cfg2_tspan(10)
    sol, st, _, rc = xo.dopri5_mlp(om, y0, t)
    assert rc == 0
    gy = np.zeros_like(sol)
    gy[-1] = np.sign(sol[-1]) / sol[-1].size
    gp, a0, sta, _, rc = xo.dopri5_mlp_adjoint(om, t, sol, gy)
    assert rc == 0
    g.update(dopri5_sol=sol, dopri5_attempts=st.n_attempts.astype(np.int64), dopri5_accepted=st.n_accepted.astype(np.int64),
             adjoint_gparams=gp, adjoint_a0=a0, adjoint_attempts=sta.n_attempts.astype(np.int64))
    solb, stb, logb, rc = xo.dopri5_mlp(om, y0, t, controller="batch")
    g.update(dopri5_batch_sol=solb, dopri5_batch_dt=logb.dt.copy(), dopri5_batch_accepted=logb.accepted.copy())
    # the other tableaux
    for m, rt in (("bosh3", 1e-6), ("fehlberg2", 1e-4), ("adaptive_heun", 1e-4), ("dopri8", 1e-7)):
        s, stt, _, rc = xo.adaptive_rk_mlp(m, om, y0, np.linspace(0, 1, 6).astype(f32), rtol=rt, atol=rt * 1e-2)
        assert rc == 0
        g[m + "_sol"], g[m + "_attempts"] = s, stt.n_attempts.astype(np.int64)
    s, stt, _, rc = xo.adaptive_rk_mlp("dopri5", om, y0, np.linspace(0, 1.5, 5).astype(f32), rtol=1e-5, atol=1e-7,
                                       step_t=[0.33, 0.9, -1.0, 0.05, 7.0], jump_t=[0.5, 1.2, 0.051])
    g["dopri5_grid_sol"], g["dopri5_grid_nfe"] = s, stt.nfe.astype(np.int64)
    # fixed grid, small and large states
    tf = np.linspace(0, 1, 9).astype(f32)
    for m in ("euler", "midpoint", "rk4"):
        g["fixed_small_" + m] = xo.fixed_mlp(m, om, y0, tf)
    w3 = fanin_weights(64, 256, seed=1)
    y3 = np.random.default_rng(1).uniform(-1, 1, (96, 64)).astype(f32)
    g["fixed_cfg3_rk4"] = xo.fixed_mlp("rk4", xo.MLP(*w3, pre="id"), y3, np.linspace(0, 1, 11).astype(f32))
    # sde (cfg4 shapes) on supplied increments
    f4, g4 = xo.MLP(*fanin_weights(32, 64, seed=2), pre="cube"), xo.MLP(*fanin_weights(32, 64, seed=3), pre="square")
    rng = np.random.default_rng(2)
    y4 = rng.uniform(-1, 1, (80, 32)).astype(f32)
    t4 = np.linspace(0, 1, 17).astype(f32)
    dW = (np.sqrt(1 / 16) * rng.standard_normal((16, 80, 32))).astype(f32)
    g["sde_cfg4_em"] = xo.sde_mlp("em", f4, g4, y4, t4, dW)
    # history gather (cfg5 shapes, 2 batches)
    rng = np.random.default_rng(5)
    his = rng.uniform(-1, 1, (2, 307, 288, 3)).astype(f32)
    lags = (np.arange(12) + rng.uniform(0, 1, 12)).astype(f32)
    for kind in ("linear", "cubic", "bez"):
        v, d = xo.history_gather(kind, his, np.arange(288, dtype=f32), lags)
        g["gather_" + kind + "_val"], g["gather_" + kind + "_der"] = v[:, ::16], d[:, ::16]  # every 16th node
    g["gather_lags"] = lags
    return g


def large_inputs():
    """Seeded inputs of oracle_vectors_large.npz (shared with tests/test_golden.py)."""
    w3 = fanin_weights(64, 256, seed=1)
    y3 = np.random.default_rng(7).uniform(-1, 1, (40, 64)).astype(f32)
    w4 = fanin_weights(32, 64, seed=2)
    y4 = np.random.default_rng(8).uniform(-1, 1, (50, 32)).astype(f32)
    return w3, y3, np.linspace(0, 1, 6).astype(f32), w4, y4, np.linspace(0, 1.2, 5).astype(f32)


def oracle_vectors_large():
    """Adaptive solvers on large states (added with csrc/xde_tile_adaptive.cuh): cfg3's field under Dopri5, the cfg4
    drift under Bosh3 with forced grid points."""
    w3, y3, t3, w4, y4, t4 = large_inputs()
    g = {}
    s, st, _, rc = xo.dopri5_mlp(xo.MLP(*w3, pre="id"), y3, t3, rtol=1e-5, atol=1e-7)
    assert rc == 0
    g.update(dopri5_cfg3_sol=s, dopri5_cfg3_attempts=st.n_attempts.astype(np.int64), dopri5_cfg3_nfe=st.nfe.astype(np.int64))
    s, st, _, rc = xo.adaptive_rk_mlp("bosh3", xo.MLP(*w4, pre="cube"), y4, t4, rtol=1e-5, atol=1e-7, step_t=[0.25, 0.7],
                                      jump_t=[0.5])
    assert rc == 0
    g.update(bosh3_cfg4_grid_sol=s, bosh3_cfg4_grid_attempts=st.n_attempts.astype(np.int64),
             bosh3_cfg4_grid_nfe=st.nfe.astype(np.int64))
    return g


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--large-only" not in sys.argv:  # the first two files are frozen since r1h; re-create them only on purpose
        np.savez_compressed(os.path.join(OUT, "reference_fixtures.npz"), **reference_fixtures())
        np.savez_compressed(os.path.join(OUT, "oracle_vectors.npz"), **oracle_vectors())
    np.savez_compressed(os.path.join(OUT, "oracle_vectors_large.npz"), **oracle_vectors_large())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
