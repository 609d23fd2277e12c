"""GPU parity tests added in round 2: grad_t_span, the D = 5/6/7 adjoint, the adjoint at BASELINE size, the
autograd surface of ddeint, torch `Sequential` fields, shape checks of the fixed solvers, default grids of the
interpolants.  Same bar as tests/test_gpu_parity.py: states bit-exact against the oracle, batch-summed
gradients at rtol 1e-5."""
import warnings

import numpy as np
import pytest

from tests.problems import cfg2_tspan, cfg2_y0, fanin_weights, spiral_weights

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def px():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import paddlexde_b200 as px

    px._lib.lib()
    return px


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def loss_grad(sol_np):
    gy = np.zeros_like(sol_np)
    gy[-1] = np.sign(sol_np[-1]) / sol_np[-1].size
    return gy


# ------------------------------------------------------------------------------------------------
# adjoint: the shapes round 1 left out, BASELINE size, grad_t_span
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,pre,B", [(5, 24, "id", 70), (6, 31, "cube", 45), (7, 40, "square", 33), (5, 64, "id", 129)])
def test_adjoint_parity_d567(px, torch, oracle, d, h, pre, B):
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = fanin_weights(d, h, seed=h)
    field, om = px.MLPField(*w, pre=pre), oracle.MLP(*w, pre=pre)
    y0 = np.random.default_rng(3).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 5).astype(f32)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t)
    gy = loss_grad(ref)
    gy[2] = 0.01 * np.random.default_rng(4).standard_normal(gy[2].shape).astype(f32)
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=512)
    g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy)
    assert rc == 0
    s = stats.read()
    assert s.status == 0 and s.n_attempts == int(st_ref.n_attempts.sum()) and s.nfe == int(st_ref.nfe.sum())
    assert np.array_equal(a0.cpu().numpy(), a_ref)
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())
    rec, cnt = log.read()
    for b in (0, B - 1):
        _, _, _, lg, _ = oracle.dopri5_mlp_adjoint(om, t, ref, gy, log_traj=b)
        r = rec[b, :cnt[b]]
        assert cnt[b] == len(lg) and np.array_equal(r.dt, lg.dt) and np.array_equal(r.ratio, lg.ratio)


def test_adjoint_full_size_parity(px, torch, oracle):
    """cfg2 at BASELINE size (B = 2^20), the adjoint -- 87 % of the timed step.  dL/dy0 of a 4096-row random subset
    must equal the oracle run on that subset alone bit for bit (trajectories are independent under the
    per-trajectory controller), and the parameter gradients of the WHOLE batch (grid-wide queue, fp64 accumulator,
    the fold of all CTAs) must agree with the oracle's over the whole batch at rtol 1e-5."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = spiral_weights()
    field, om = px.MLPField(*w, pre="cube"), oracle.MLP(*w, pre="cube")
    B = 1 << 20
    y0, t = cfg2_y0(B), cfg2_tspan(10)
    xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), t)
    s = px.Dopri5(xde=xde, y0=xde.y0, rtol=1e-7, atol=1e-9, controller="trajectory")
    sol = s.integrate(t)
    gy = torch.zeros_like(sol)
    gy[-1] = torch.sign(sol[-1]) / sol[-1].numel()
    g, a0, stats, _ = adjoint_backward(field, t, sol, gy, return_adj_y0=True)
    sol_h, gy_h = sol.cpu().numpy(), gy.cpu().numpy()
    g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, sol_h, gy_h)   # ~15 s on 16 cores
    assert rc == 0
    st = stats.read()
    assert st.status == 0
    assert st.n_attempts == int(st_ref.n_attempts.sum()) and st.n_accepted == int(st_ref.n_accepted.sum())
    idx = np.random.default_rng(11).choice(B, 4096, replace=False)
    assert np.array_equal(a0.cpu().numpy()[idx], a_ref[idx])
    # the subset alone reproduces the same rows (independence of the trajectories)
    _, a_sub, _, _, _ = oracle.dopri5_mlp_adjoint(om, t, sol_h[:, idx], gy_h[:, idx])
    assert np.array_equal(a_sub, a_ref[idx])
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())


@pytest.mark.parametrize("d,h,pre,B,reverse", [(2, 50, "cube", 300, False), (4, 32, "id", 65, False),
                                               (3, 17, "square", 40, True), (1, 16, "id", 33, False)])
def test_adjoint_grad_t_span_matches_oracle(px, torch, oracle, d, h, pre, B, reverse):
    """t_requires_grad branch (functional/odeint_adjoint.py:129-141,161-162): grad_t_span[i] = f(t_i, y_i).grad_y[i],
    grad_t_span[0] = the final aug_state[0]; the g_t slot takes part in the controller norm, so the step sequence
    differs from the run without it -- and still equals the oracle's."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = spiral_weights() if (d, h) == (2, 50) else fanin_weights(d, h, seed=h)
    field, om = px.MLPField(*w, pre=pre), oracle.MLP(*w, pre=pre)
    y0 = cfg2_y0(B) if d == 2 else np.random.default_rng(3).uniform(-1, 1, (B, d)).astype(f32)
    t = cfg2_tspan(6) if d == 2 else np.linspace(0, 1, 5).astype(f32)
    if reverse:
        t = t[::-1].copy()
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t)
    gy = (np.random.default_rng(5).standard_normal(ref.shape) / ref[0].size).astype(f32)
    gt = torch.full((t.size,), float("nan"), device="cuda")
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=512, out_grad_t=gt)
    gt_ref = np.zeros(t.size, f32)
    g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, grad_t=gt_ref)
    assert rc == 0
    s = stats.read()
    assert s.status == 0 and s.n_attempts == int(st_ref.n_attempts.sum())
    assert np.array_equal(a0.cpu().numpy(), a_ref)
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())
    np.testing.assert_allclose(gt.cpu().numpy(), gt_ref, rtol=1e-5, atol=1e-6 * np.abs(gt_ref).max())
    rec, cnt = log.read()
    for b in (0, B - 1):
        _, _, _, lg, _ = oracle.dopri5_mlp_adjoint(om, t, ref, gy, log_traj=b, grad_t=np.zeros(t.size, f32))
        r = rec[b, :cnt[b]]
        assert cnt[b] == len(lg) and np.array_equal(r.dt, lg.dt) and np.array_equal(r.ratio, lg.ratio)
    # the analytic value for an autonomous field: dL/dt_i = f(y_i).gy_i (i >= 1), dL/dt_0 = -sum of those
    f_all = om(0.0, ref.reshape(-1, d)).reshape(ref.shape)
    dl = (f_all.astype(np.float64) * gy).sum(axis=(1, 2))
    want = np.concatenate([[-dl[1:].sum()], dl[1:]])
    np.testing.assert_allclose(gt.cpu().numpy(), want, rtol=2e-4, atol=2e-5 * np.abs(want).max())


def test_odeint_adjoint_returns_grad_t_span(px, torch, oracle):
    w = [torch.tensor(a, device="cuda", requires_grad=True) for a in spiral_weights()]
    field = px.MLPField(*w, pre="cube")
    y0 = torch.from_numpy(cfg2_y0(64)).cuda()
    t = torch.tensor(cfg2_tspan(5), requires_grad=True)
    sol = px.odeint_adjoint(field, y0, t, solver=px.Dopri5, options={"controller": "trajectory"})
    (sol ** 2).sum().backward()
    assert t.grad is not None and t.grad.shape == t.shape and torch.isfinite(t.grad).all()
    assert abs(float(t.grad.sum())) <= 1e-4 * float(t.grad.abs().max())  # time-shift invariance of an autonomous field
    assert all(p.grad is not None for p in w)
    with pytest.raises(NotImplementedError):  # the batch-controller adjoint has no g_t slot
        sol = px.odeint_adjoint(field, y0, t, solver=px.Dopri5, options={"controller": "batch"})
        (sol ** 2).sum().backward()


def test_default_controller_is_announced_once(px, torch, oracle):
    import paddlexde_b200.solver.adaptive_solver as A

    field = px.MLPField(*spiral_weights(), pre="cube")
    y0 = torch.from_numpy(cfg2_y0(8)).cuda()
    A._warned_default = False
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        px.odeint(field, y0, cfg2_tspan(3), px.Dopri5)
        px.odeint(field, y0, cfg2_tspan(3), px.Dopri5)
        px.odeint(field, y0, cfg2_tspan(3), px.Dopri5, options={"controller": "batch"})
    assert sum(issubclass(r.category, A.ControllerDefaultWarning) for r in rec) == 1


# ------------------------------------------------------------------------------------------------
# ddeint autograd, torch Sequential fields, shape checks, default interpolant grid
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["Euler", "Midpoint", "RK4"])
def test_ddeint_solution_is_differentiable(px, torch, oracle, solver):
    """The D3STN trainer backpropagates loss(preds) with preds = ddeint(...)[0] (example/D3STN/train_dde.py:424-454):
    the gradient must reach func's parameters through the damped fuse and the lags through HistoryIndex.  Checked
    against torch autograd through the same recursion written with torch ops."""
    rng = np.random.default_rng(6)
    R, N, Th, D, L = 3, 11, 40, 3, 6
    his = torch.tensor(rng.uniform(-1, 1, (R, N, Th, D)).astype(f32), device="cuda")
    span = torch.arange(Th, dtype=torch.float32, device="cuda")
    lags0 = (np.arange(L) * 3 + rng.uniform(0.1, 0.9, L)).astype(f32)
    y0 = torch.tensor(rng.uniform(-1, 1, (R, N, L, D)).astype(f32), device="cuda")
    t = np.array([0.0, 0.5, 1.25, 2.0], f32)
    V0 = rng.standard_normal((D, D)).astype(f32) * 0.3
    res = {}
    for mode in ("kernel", "torch"):
        lags = torch.tensor(lags0, device="cuda", requires_grad=True)
        V = torch.tensor(V0, device="cuda", requires_grad=True)

        def func(y_lags, y):
            return torch.tanh(y @ V + y_lags.mean(dim=-2, keepdim=True))

        if mode == "kernel":
            sol, y_lags = px.ddeint(func, y0, t, lags, his, span, getattr(px, solver))
        else:
            y_lags = px.xde.base_dde.HistoryIndex.apply(lags, his, span)  # same gather; the recursion in torch ops

            def fuse(dy, dt, yy):
                return (dy - 0.001 * (dy * dt + yy)) * dt + yy

            class X:
                move = staticmethod(lambda t0, dt, yy: func(y_lags, yy))
            X.fuse = staticmethod(fuse)
            from paddlexde_b200.functional.ddeint import _step
            y, sols = y0, [y0]
            for i in range(1, t.size):
                y, _ = _step(getattr(px, solver).method, X, float(t[i - 1]), float(t[i]), y)
                sols.append(y)
            sol = torch.cat(sols, dim=-2)
        assert sol.requires_grad and tuple(sol.shape) == (R, N, L * t.size, D)
        wgt = torch.linspace(0.5, 1.5, sol.numel(), device="cuda").reshape(sol.shape)
        (sol * wgt).sum().backward()
        res[mode] = (sol.detach().cpu().numpy(), V.grad.cpu().numpy(), lags.grad.cpu().numpy())
    np.testing.assert_allclose(res["kernel"][0], res["torch"][0], rtol=2e-6, atol=2e-6)
    for k in (1, 2):
        assert np.abs(res["torch"][k]).max() > 0
        np.testing.assert_allclose(res["kernel"][k], res["torch"][k], rtol=2e-5, atol=2e-5 * np.abs(res["torch"][k]).max())


def test_torch_sequential_field_trains_all_parameters(px, torch, oracle):
    """MLPField.from_sequential keeps the module's own Parameters as the leaves ([out, in] layout): weights AND biases
    receive gradients (transposed back), and refresh() sees an optimizer step."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, 50), torch.nn.Tanh(), torch.nn.Linear(50, 2)).cuda()
    field = px.MLPField.from_sequential(net, pre="cube")
    assert [p is q for p, q in zip(field.parameters(), net.parameters())] == [True] * 4
    y0 = torch.from_numpy(cfg2_y0(128)).cuda()
    t = cfg2_tspan(5)
    opt = torch.optim.SGD(net.parameters(), lr=0.05)
    sol = px.odeint_adjoint(field, y0, t, solver=px.Dopri5, options={"controller": "trajectory"})
    loss = sol[-1].abs().mean()
    loss.backward()
    grads = [p.grad.clone() for p in net.parameters()]
    assert all(g is not None and torch.isfinite(g).all() and g.abs().max() > 0 for g in grads)
    # against the oracle with the weights in [in, out] layout
    w = [net[0].weight.detach().t().contiguous().cpu().numpy(), net[0].bias.detach().cpu().numpy(),
         net[2].weight.detach().t().contiguous().cpu().numpy(), net[2].bias.detach().cpu().numpy()]
    om = oracle.MLP(*w, pre="cube")
    ref, _, _, _ = oracle.dopri5_mlp(om, y0.cpu().numpy(), t)
    assert np.array_equal(sol.detach().cpu().numpy(), ref)
    g_ref, _, _, _, _ = oracle.dopri5_mlp_adjoint(om, t, ref, loss_grad(ref))
    gw1, gb1, gw2, gb2 = om.split(g_ref)
    for got, want in zip(grads, (gw1.T, gb1, gw2.T, gb2)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())
    opt.step()
    field.refresh()
    sol2 = px.odeint(field, y0, t, px.Dopri5, options={"controller": "trajectory"})
    assert not np.array_equal(sol2.cpu().numpy(), ref), "refresh() must re-read the stepped parameters"


def test_fixed_solvers_refuse_a_state_dim_mismatch(px, torch, oracle):
    field = px.MLPField(*fanin_weights(4, 32), pre="id")
    y0 = torch.zeros(16, 1, 3, device="cuda")
    with pytest.raises(ValueError, match="state dim"):
        px.odeint(field, y0, np.linspace(0, 1, 4).astype(f32), px.RK4)
    with pytest.raises(ValueError, match="state dim"):
        px.sdeint(field, field, y0, np.linspace(0, 1, 4).astype(f32), px.Euler,
                  options={"bm_increments": torch.zeros(3, 16, 3, device="cuda")})


@pytest.mark.parametrize("cls,kind", [("LinearInterpolation", "linear"), ("CubicHermiteSpline", "cubic"), ("BezierSpline", "bez")])
def test_interpolants_default_grid(px, torch, oracle, cls, kind):
    """t=None: the grid is 0..n-1 (interpolate_base.py:21-27 reads the first n points of its linspace)."""
    rng = np.random.default_rng(2)
    series = rng.standard_normal((5, 12, 3)).astype(f32)
    q = np.array([0.0, 0.4, 3.5, 10.25, 11.0], f32)
    it = getattr(px.interpolation, cls)(torch.from_numpy(series).cuda())
    val, der = it.evaluate(q), it.derivative(q)
    v_ref, d_ref = oracle.history_gather(kind, series, np.arange(12, dtype=f32), q)
    assert np.array_equal(val.cpu().numpy(), v_ref) and np.array_equal(der.cpu().numpy(), d_ref)


# ------------------------------------------------------------------------------------------------
# adjoint for large states (csrc/xde_adj_tile.cu): tiles of trajectories per CTA, gradients in tensor memory
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,pre,B", [(64, 256, "id", 70), (64, 128, "cube", 33), (32, 256, "id", 65), (32, 128, "square", 130),
                                        (32, 64, "cube", 129), (16, 64, "id", 200)])
def test_adjoint_large_state_parity(px, torch, oracle, d, h, pre, B):
    """odeint_adjoint's backward for D = 16 / 32 / 64 (cfg3's 64-256-64 field): dL/dy0 bit-exact, the per-trajectory
    (dt, ratio, accept) sequences identical, parameter gradients within rtol 1e-5 of the oracle."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = fanin_weights(d, h, seed=h + d)
    field, om = px.MLPField(*w, pre=pre), oracle.MLP(*w, pre=pre)
    y0 = np.random.default_rng(3).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 4).astype(f32)
    kw = dict(rtol=1e-6, atol=1e-8)
    ref, _, _, rc = oracle.dopri5_mlp(om, y0, t, **kw)
    assert rc == 0
    gy = loss_grad(ref)
    gy[1] = (0.01 / ref[1].size) * np.random.default_rng(4).standard_normal(gy[1].shape).astype(f32)
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=256, **kw)
    g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, **kw)
    assert rc == 0
    s = stats.read()
    assert s.status == 0
    assert s.n_attempts == int(st_ref.n_attempts.sum()) and s.n_accepted == int(st_ref.n_accepted.sum())
    assert s.nfe == int(st_ref.nfe.sum())
    assert np.array_equal(a0.cpu().numpy(), a_ref), "adjoint state dL/dy0 must be bit-exact"
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())
    rec, cnt = log.read()
    for b in (0, B // 2, B - 1):
        _, _, _, lg, _ = oracle.dopri5_mlp_adjoint(om, t, ref, gy, log_traj=b, **kw)
        assert cnt[b] == len(lg)
        r = rec[b, :cnt[b]]
        assert np.array_equal(r.accepted, lg.accepted) and np.array_equal(r.dt, lg.dt) and np.array_equal(r.ratio, lg.ratio)


def test_adjoint_large_state_rejections_reverse_time_and_status(px, torch, oracle):
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    d, h, B = 32, 64, 97
    w = [2.5 * a for a in fanin_weights(d, h, seed=9)]
    field, om = px.MLPField(*w, pre="id"), oracle.MLP(*w, pre="id")
    y0 = np.random.default_rng(1).uniform(-1, 1, (B, d)).astype(f32)
    kw = dict(rtol=1e-5, atol=1e-7)
    for t in (np.linspace(0, 1.5, 4).astype(f32), np.linspace(1.5, 0, 4).astype(f32)):
        ref, _, _, rc = oracle.dopri5_mlp(om, y0, t, **kw)
        assert rc == 0
        gy = (np.random.default_rng(5).standard_normal(ref.shape) / ref[0].size).astype(f32)
        g, a0, stats, _ = adjoint_backward(field, t, ref, gy, return_adj_y0=True, **kw)
        g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, **kw)
        assert rc == 0
        if t[1] > t[0]:
            assert (st_ref.n_attempts > st_ref.n_accepted).any(), "the case must exercise rejections (replay passes)"
        s = stats.read()
        assert s.status == 0 and s.n_attempts == int(st_ref.n_attempts.sum()) and s.n_accepted == int(st_ref.n_accepted.sum())
        assert np.array_equal(a0.cpu().numpy(), a_ref)
        np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=2e-6 * np.abs(g_ref).max())
    # max_num_steps -> status word, NaN rows
    t = np.array([0.0, 3.0], f32)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, **kw)
    gy = np.ones_like(ref) / ref[0].size
    with pytest.raises(AssertionError, match="max_num_steps"):
        adjoint_backward(field, t, ref, gy, max_num_steps=2, **kw)


def test_large_state_neural_ode_trains_through_the_public_api(px, torch, oracle):
    """A 64-256-64 neural ODE through odeint_adjoint + backward(): gradients of every parameter against fp64 torch
    autograd through a fine RK4 integration (the true gradient, rtol 2e-3 at these tolerances)."""
    rng = np.random.default_rng(0)
    d, h, B = 64, 256, 48
    w = fanin_weights(d, h, seed=3)
    tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in w]
    field = px.MLPField(*tw, pre="id")
    y0 = torch.tensor(rng.uniform(-1, 1, (B, d)).astype(f32), device="cuda")
    t = np.linspace(0, 0.5, 3).astype(f32)
    target = torch.tensor(rng.uniform(-1, 1, (B, d)).astype(f32), device="cuda")
    sol = px.odeint_adjoint(field, y0, t, solver=px.Dopri5, rtol=1e-6, atol=1e-8, options={"controller": "trajectory"})
    loss = ((sol[-1] - target) ** 2).mean() + 0.1 * (sol[1] ** 2).mean()
    loss.backward()
    got = [p.grad.double().cpu() for p in tw]
    W = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in w]
    f = lambda y: torch.tanh(y @ W[0] + W[1]) @ W[2] + W[3]
    y = y0.double().cpu()
    outs = [y]
    for i in range(1, t.size):
        n = 40
        hh = float(t[i] - t[i - 1]) / n
        for _ in range(n):
            k1 = f(y); k2 = f(y + 0.5 * hh * k1); k3 = f(y + 0.5 * hh * k2); k4 = f(y + hh * k3)
            y = y + hh / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        outs.append(y)
    ref_loss = ((outs[-1] - target.double().cpu()) ** 2).mean() + 0.1 * (outs[1] ** 2).mean()
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    for gq, Wq in zip(got, W):
        np.testing.assert_allclose(gq.numpy(), Wq.grad.numpy(), rtol=2e-3, atol=2e-4 * float(Wq.grad.abs().max()))


# ------------------------------------------------------------------------------------------------
# Milstein for large states (north_star (4) at cfg4's shape): FP32 tiles, bit-exact against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,B,pref,preg", [(32, 64, 130, "cube", "square"), (64, 128, 33, "id", "id"), (16, 64, 70, "id", "cube"),
                                              (32, 128, 65, "square", "id")])
def test_milstein_large_state_bit_exact(px, torch, oracle, d, h, B, pref, preg):
    rng = np.random.default_rng(d + h)
    wf, wg = fanin_weights(d, h, seed=2), fanin_weights(d, h, seed=3)
    f, g = px.MLPField(*wf, pre=pref), px.MLPField(*wg, pre=preg)
    y0 = rng.uniform(-1, 1, (B, 1, d)).astype(f32)
    t = np.linspace(0, 1, 9).astype(f32)
    dW = (0.3 * rng.standard_normal((8, B, d))).astype(f32)
    ref = oracle.sde_mlp("milstein", oracle.MLP(*wf, pref), oracle.MLP(*wg, preg), y0[:, 0], t, dW)
    got = px.sdeint(f, g, torch.from_numpy(y0).cuda(), t, px.Euler,
                    options={"bm_increments": torch.from_numpy(dW).cuda(), "scheme": "milstein"})  # math="auto" -> FP32 tiles
    assert np.array_equal(got.cpu().numpy(), ref)
    assert not np.array_equal(ref, oracle.sde_mlp("em", oracle.MLP(*wf, pref), oracle.MLP(*wg, preg), y0[:, 0], t, dW))
    # the device-side generator gives the same solution as its own table
    tab = px.utils.brownian.brownian_increments(7, t, B, d)
    a = px.sdeint(f, g, torch.from_numpy(y0).cuda(), t, px.Euler, options={"bm_seed": 7, "scheme": "milstein"})
    b = px.sdeint(f, g, torch.from_numpy(y0).cuda(), t, px.Euler, options={"bm_increments": tab, "scheme": "milstein"})
    assert torch.equal(a, b)
    with pytest.raises(px.UnsupportedFieldError):  # the tcgen05 path has no Milstein: no silent change of arithmetic
        px.sdeint(f, g, torch.from_numpy(y0).cuda(), t, px.Euler,
                  options={"bm_increments": torch.from_numpy(dW).cuda(), "scheme": "milstein", "math": "tensor"})


# ------------------------------------------------------------------------------------------------
# FixedSolver(step_size= | grid_constructor=): the reference's loop as it is (base_fixed_solver.py:49-89,119-139)
# ------------------------------------------------------------------------------------------------
def _reference_grid_solution(oracle, method, om, y0, t, grid):
    """The reference's FixedSolver.integrate restated: len(t) - 1 steps on grid[0..len(t)), output i = linear_interp of
    step i evaluated at t[i] (interp_fn.py:4-10), fp32 throughout."""
    g = np.ascontiguousarray(grid[:t.size])
    y = oracle.fixed_mlp(method, om, y0, g)            # [B, T, D] on the grid
    out = np.empty_like(y)
    out[:, 0] = y[:, 0]
    for i in range(1, t.size):
        t0, t1, tq = g[i - 1], g[i], t[i]
        if tq == t0:
            out[:, i] = y[:, i - 1]
        elif tq == t1:
            out[:, i] = y[:, i]
        else:
            slope = np.float32(np.float32(tq - t0) / np.float32(t1 - t0))
            out[:, i] = y[:, i - 1] + slope * (y[:, i] - y[:, i - 1])
    return out


@pytest.mark.parametrize("solver", ["Euler", "RK4", "Midpoint"])
@pytest.mark.parametrize("d,h", [(2, 50), (32, 64)])
def test_fixed_solver_step_size_and_grid_constructor(px, torch, oracle, solver, d, h):
    w = spiral_weights() if d == 2 else fanin_weights(d, h, seed=4)
    pre = "cube" if d == 2 else "id"
    field, om = px.MLPField(*w, pre=pre), oracle.MLP(*w, pre=pre)
    B = 77
    y0 = np.random.default_rng(2).uniform(-1, 1, (B, 1, d)).astype(f32)
    t = np.linspace(0, 1, 6).astype(f32)
    step = 0.05
    niters = int(np.ceil(np.float32((t[-1] - t[0]) / np.float32(step)) + np.float32(1.0)))
    grid = np.arange(0, niters, dtype=f32) * np.float32(step) + t[0]
    grid[-1] = t[-1]
    ref = _reference_grid_solution(oracle, solver.lower(), om, y0[:, 0], t, grid)
    got = px.odeint(field, torch.from_numpy(y0).cuda(), t, getattr(px, solver), options={"step_size": step, "math": "fp32"})
    assert tuple(got.shape) == (B, t.size, d) and np.array_equal(got.cpu().numpy(), ref)
    # the same grid through grid_constructor; the plain call (grid == t_span) differs
    got2 = px.odeint(field, torch.from_numpy(y0).cuda(), t, getattr(px, solver),
                     options={"grid_constructor": lambda y, tt: grid, "math": "fp32"})
    assert torch.equal(got, got2)
    plain = px.odeint(field, torch.from_numpy(y0).cuda(), t, getattr(px, solver), options={"math": "fp32"})
    assert not torch.equal(plain, got)
    with pytest.raises(ValueError):
        px.odeint(field, torch.from_numpy(y0).cuda(), t, px.RK4, options={"step_size": 0.1, "grid_constructor": lambda y, tt: grid})
    with pytest.raises(AssertionError):  # the reference asserts the grid's end points
        px.odeint(field, torch.from_numpy(y0).cuda(), t, px.RK4, options={"grid_constructor": lambda y, tt: grid[:-1]})
    with pytest.raises(NotImplementedError):
        px.odeint(field, torch.from_numpy(y0).cuda(), t, px.RK4, options={"step_size": step, "interp": "cubic"})
