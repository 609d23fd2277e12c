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


# ------------------------------------------------------------------------------------------------
# odeint_adjoint's backward against vectors produced by the reference's own functional/odeint_adjoint.py
# (tools/make_reference_adjoint_golden.py; CPU twin: tests/test_reference_run_adjoint_golden.py)
# ------------------------------------------------------------------------------------------------
ZA = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_adjoint_vectors.npz"), allow_pickle=False)
# the reference's controller = the batch-controller kernel (D in {1, 2, 4}, H <= 64, no g_t slot): everything bit for bit
ADJ_BATCH = ["b1_cfg2_default_mixed", "b1_cfg2_seminorm", "b1_d1_square", "batch_cfg1_default_mixed", "batch_B70_rejections",
             "batch_d4_seminorm_rejections", "batch_d4_options", "batch_d1"]
# B = 1: the reference's controller = the per-trajectory kernels (seminorm; D <= 8 with grad_t_span, D >= 16 on tiles)
ADJ_B1 = ["b1_cfg2_seminorm", "b1_cfg2_grad_t", "b1_d4_options", "b1_d3", "b1_d8", "b1_d5_rejections_seminorm",
          "b1_d2_reverse_span", "b1_D32", "b1_D64"]


def adjoint_case(px, name, controller):
    import torch
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    meta = ast.literal_eval(str(ZA[f"{name}/meta"]))
    field = px.MLPField(ZA[f"{name}/w1"], ZA[f"{name}/b1"], ZA[f"{name}/w2"], ZA[f"{name}/b2"], pre=meta["pre"])
    kw = {k: v for k, v in meta.items() if k not in ("pre", "adj_norm")}
    t = ZA[f"{name}/t"]
    want_gt = f"{name}/grad_t" in ZA.files
    gt = torch.full((t.size,), float("nan"), device="cuda") if want_gt else None
    g, a0, stats, log = adjoint_backward(field, t, ZA[f"{name}/sol"], ZA[f"{name}/grad_y"], return_adj_y0=True,
                                         log_attempts=512, controller=controller, adj_norm=meta["adj_norm"],
                                         out_grad_t=gt, **{"rtol": 1e-7, "atol": 1e-9, **kw})
    assert stats.read().status == 0
    g_ref = np.concatenate([ZA[f"{name}/{k}"].ravel() for k in ("gw1", "gb1", "gw2", "gb2")])
    rec, cnt = log.read()
    rlog = ZA[f"{name}/log"]
    assert int(cnt[0]) == len(rlog)
    r = rec[0, :len(rlog)]
    for f in ("t0", "dt", "ratio", "accepted"):
        assert np.array_equal(r[f], rlog[f]), f"{name}: backward attempt log field {f}"
    assert np.array_equal(a0.cpu().numpy(), ZA[f"{name}/adj_y0"]), f"{name}: dL/dy0"
    return g.cpu().numpy(), g_ref, (None if gt is None else gt.cpu().numpy())


@pytest.mark.parametrize("name", ADJ_BATCH)
def test_batch_adjoint_kernel_reproduces_the_reference_run(px, name):
    """The reference's default configuration (global controller; mixed norm unless 'seminorm'): attempt log, dL/dy0 and
    the parameter gradients (a state of this solve) bit for bit."""
    g, g_ref, _ = adjoint_case(px, name, "batch")
    assert np.array_equal(g, g_ref), f"{name}: max|d| = {np.abs(g - g_ref).max()}"


@pytest.mark.parametrize("name", ADJ_B1)
def test_per_trajectory_adjoint_kernels_reproduce_the_reference_run_at_b1(px, name):
    """Attempt log and dL/dy0 bit for bit; the parameter gradients and grad_t_span are folded outside the Runge-Kutta
    state by these kernels (same contributions, another fp32 order): within 3e-5 of the largest entry under the
    vectors' random-sign cotangents (the bound the fuzz sweep measured: 1.2e-5)."""
    g, g_ref, gt = adjoint_case(px, name, "trajectory")
    np.testing.assert_allclose(g, g_ref, rtol=1e-5, atol=3e-5 * np.abs(g_ref).max())
    if gt is not None:
        gt_ref = ZA[f"{name}/grad_t"]
        np.testing.assert_allclose(gt, gt_ref, rtol=1e-5, atol=3e-5 * np.abs(gt_ref).max())


@pytest.mark.parametrize("name", ["b1_cfg2_default_mixed", "batch_cfg1_default_mixed", "batch_d4_options",
                                  "batch_d4_seminorm_rejections"])
def test_odeint_adjoint_entry_point_reproduces_the_reference_run(px, name):
    """The same vectors through the PUBLIC entry point, called the way the reference was called by the generator:
    `odeint_adjoint(func, y0, t, rtol=, atol=, solver=Dopri5, options={"norm": _rms_norm, **solver options},
    adjoint_options=None | {"norm": "seminorm", ...})` then `sol.backward(grad_y)` -- the option defaulting of
    functional/odeint_adjoint.py:194-238 (backward tolerances and solver options inherited from the forward ones, mixed
    norm unless "seminorm") has to route everything to the two kernels.  `controller="batch"` = the reference's."""
    import torch

    meta = ast.literal_eval(str(ZA[f"{name}/meta"]))
    pre, adj_norm = meta.pop("pre"), meta.pop("adj_norm")
    rtol, atol = meta.pop("rtol", 1e-7), meta.pop("atol", 1e-9)
    params = [torch.tensor(ZA[f"{name}/{k}"], device="cuda", requires_grad=True) for k in ("w1", "b1", "w2", "b2")]
    field = px.MLPField(*params, pre=pre)
    sol = px.odeint_adjoint(field, torch.from_numpy(ZA[f"{name}/y0"]).cuda(), ZA[f"{name}/t"], rtol=rtol, atol=atol,
                            solver=px.Dopri5, options={"norm": px.utils._rms_norm, "controller": "batch", **meta},
                            adjoint_options=({"norm": "seminorm", "controller": "batch", **meta} if adj_norm == "seminorm"
                                             else None))
    assert np.array_equal(sol.detach().cpu().numpy(), ZA[f"{name}/sol"])
    sol.backward(torch.from_numpy(ZA[f"{name}/grad_y"]).cuda())
    assert px.odeint_adjoint.last["adj_norm"] == adj_norm
    for p, k in zip(params, ("gw1", "gb1", "gw2", "gb2")):
        assert np.array_equal(p.grad.cpu().numpy(), ZA[f"{name}/{k}"]), k


# ------------------------------------------------------------------------------------------------
# the delay path against vectors produced by the reference's own interpolation / base_dde / ddeint code
# (tools/make_reference_dde_golden.py; CPU twin: tests/test_reference_run_dde_golden.py)
# ------------------------------------------------------------------------------------------------
ZD = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_dde_vectors.npz"), allow_pickle=False)
GATHER = sorted({k.split("/")[1] for k in ZD.files if k.startswith("gather/")})
DDEINT = sorted({k.split("/")[1] for k in ZD.files if k.startswith("ddeint/") and k.count("/") == 2})


@pytest.mark.parametrize("kind", ["linear", "cubic", "bez"])
@pytest.mark.parametrize("name", GATHER)
def test_gather_kernel_reproduces_the_reference_interpolants(px, name, kind):
    from paddlexde_b200.xde.base_dde import history_gather

    v, d = history_gather(ZD[f"gather/{name}/lags"], ZD[f"gather/{name}/his"], ZD[f"gather/{name}/span"], kind)
    assert np.array_equal(v.cpu().numpy(), ZD[f"gather/{name}/{kind}/val"]), "evaluate"
    assert np.array_equal(d.cpu().numpy(), ZD[f"gather/{name}/{kind}/der"]), "derivative"


def test_history_index_reproduces_the_reference_forward_and_backward(px):
    import torch

    lags = torch.tensor(ZD["index/lags"], device="cuda", requires_grad=True)
    y_lags = px.xde.base_dde.HistoryIndex.apply(lags, torch.from_numpy(ZD["index/his"]).cuda(),
                                                torch.from_numpy(ZD["index/span"]).cuda())
    assert np.array_equal(y_lags.detach().cpu().numpy(), ZD["index/y_lags"])
    y_lags.backward(torch.from_numpy(ZD["index/grad_y"]).cuda())
    ref = ZD["index/grad_lags"]  # the kernel reduces in another order: rtol 1e-5 (as test_history_gather_cfg5)
    np.testing.assert_allclose(lags.grad.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


@pytest.mark.parametrize("name", DDEINT)
def test_ddeint_reproduces_the_reference_run(px, name):
    import torch
    from tests.problems import dde_field_coefficients

    ca, cb = (float(c) for c in dde_field_coefficients())
    func = lambda y_lags, y: y_lags * ca - y * cb  # noqa: E731  (elementwise fp32: the bits of the NumPy field)
    sol, y_lags = px.ddeint(func, torch.from_numpy(ZD["ddeint/y0"]).cuda(), ZD[f"ddeint/{name}/t"],
                            torch.from_numpy(ZD["index/lags"]).cuda(), torch.from_numpy(ZD["index/his"]).cuda(),
                            torch.from_numpy(ZD["index/span"]).cuda(), getattr(px, name.split("_")[0].capitalize() if
                                                                                 not name.startswith("rk4") else "RK4"),
                            fixed_solver_interp=str(ZD[f"ddeint/{name}/interp"]))
    assert np.array_equal(y_lags.cpu().numpy(), ZD["index/y_lags"])
    assert np.array_equal(sol.cpu().numpy(), ZD[f"ddeint/{name}/sol"])


# ------------------------------------------------------------------------------------------------
# the reference's assertions (tools/make_reference_error_golden.py; CPU twin: tests/test_reference_run_error_golden.py)
# ------------------------------------------------------------------------------------------------
ZE = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_errors.npz"), allow_pickle=False)
ERR_NAMES = sorted({k.split("/")[0] for k in ZE.files})
ERR_CODE = {"max_num_steps exceeded": 3, "non-finite values in state `y`": 2, "underflow in dt": 1}


@pytest.mark.parametrize("name", ERR_NAMES)
def test_kernels_raise_the_reference_assertions(px, name):
    import torch

    meta = ast.literal_eval(str(ZE[f"{name}/meta"]))
    field = px.MLPField(ZE[f"{name}/w1"], ZE[f"{name}/b1"], ZE[f"{name}/w2"], ZE[f"{name}/b2"], pre=meta.pop("pre"))
    y0, t = torch.from_numpy(ZE[f"{name}/y0"]).cuda(), ZE[f"{name}/t"]
    msg = str(ZE[f"{name}/message"])
    (prefix, code), = [(p, c) for p, c in ERR_CODE.items() if msg.startswith(p)]
    done = int(ZE[f"{name}/adaptive_step_calls"]) - (0 if code == 3 else 1)
    B = y0.shape[0]
    for controller in (("batch", "trajectory") if B == 1 else ("batch",)):  # the reference's controller is the global one
        xde = px.xde.BaseODE(field, y0, t)
        kw = {"rtol": 1e-7, "atol": 1e-9, **meta}
        with pytest.raises(AssertionError) as e:
            px.Dopri5(xde=xde, y0=xde.y0, controller=controller, **kw).integrate(t)
        assert str(e.value).startswith(prefix), (controller, str(e.value))
        s = px.Dopri5(xde=xde, y0=xde.y0, controller=controller, check_status=False, **kw)
        s.integrate(t)
        st = s.read_stats()
        assert st.status == code and st.n_attempts == done * B, (controller, st)
