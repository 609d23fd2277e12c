"""GPU parity tests: the sm_100a kernels (called through the C ABI of libxde_b200.so via the shim)
against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): identical accept/reject step sequences; final states and gradients
within rtol 1e-5 (fp32).  Because the kernels follow the oracle's arithmetic specification, states
are in fact compared BIT-EXACT here; parameter gradients (fp64 batch accumulation on both sides, in
different orders) are compared at rtol 1e-5 / atol 1e-6*max|g|.
"""
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

    px._lib.lib()  # hard failure if the extension is missing: there is no fallback to test
    return px


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def both(px, oracle, w, pre):
    return px.MLPField(*w, pre=pre), oracle.MLP(*w, pre=pre)


def solve_fwd(px, torch, field, y0, t, **opt):
    xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), t)
    s = px.Dopri5(xde=xde, y0=xde.y0, **{"rtol": 1e-7, "atol": 1e-9, **opt})
    sol = s.integrate(t)
    return sol, s


# ------------------------------------------------------------------------------------------------
# dopri5 forward
# ------------------------------------------------------------------------------------------------
def test_dopri5_cfg1_bit_exact_and_step_sequence(px, torch, oracle):
    """cfg1: spiral ODE, MLP 2-50-2, B=20, t = linspace(0,25,1000)[:32]."""
    field, om = both(px, oracle, spiral_weights(), "cube")
    rng = np.random.default_rng(42)
    y0 = (np.array([2.0, 0.0]) + rng.standard_normal((20, 2))).astype(f32)
    t = cfg2_tspan(32)
    sol, s = solve_fwd(px, torch, field, y0, t, log_attempts=256)
    ref, st, _, rc = oracle.dopri5_mlp(om, y0, t)
    assert rc == 0
    assert np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum())
    assert s.stats.n_accepted == int(st.n_accepted.sum())
    assert s.stats.nfe == int(st.nfe.sum())
    rec, cnt = s.attempt_log.read()
    for b in range(20):  # identical accept/reject sequence, dt and error ratio, per trajectory
        _, _, lg, _ = oracle.dopri5_mlp(om, y0, t, log_traj=b)
        assert cnt[b] == len(lg)
        r = rec[b, :cnt[b]]
        assert np.array_equal(r.accepted, lg.accepted)
        assert np.array_equal(r.dt, lg.dt) and np.array_equal(r.t0, lg.t0)
        assert np.array_equal(r.ratio, lg.ratio)


@pytest.mark.parametrize("d,h,pre", [(1, 16, "id"), (2, 50, "cube"), (3, 20, "square"), (4, 33, "id"), (8, 64, "id"),
                                     (5, 24, "id"), (6, 31, "cube"), (7, 40, "square")])
def test_dopri5_shapes(px, torch, oracle, d, h, pre):
    field, om = both(px, oracle, fanin_weights(d, h, seed=d), pre)
    y0 = np.random.default_rng(d).uniform(-1, 1, (333, d)).astype(f32)
    t = np.linspace(0, 2, 7).astype(f32)
    sol, s = solve_fwd(px, torch, field, y0, t, rtol=1e-6, atol=1e-8)
    ref, st, _, rc = oracle.dopri5_mlp(om, y0, t, rtol=1e-6, atol=1e-8)
    assert rc == 0 and np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum())


def test_dopri5_rejections_reverse_time_and_first_step(px, torch, oracle):
    w = [3.0 * a for a in fanin_weights(2, 50, seed=5)]
    field, om = both(px, oracle, w, "id")
    y0 = np.random.default_rng(1).uniform(-1, 1, (257, 2)).astype(f32)
    t = np.linspace(0, 4, 9).astype(f32)
    sol, s = solve_fwd(px, torch, field, y0, t, rtol=1e-6, atol=1e-8)
    ref, st, _, _ = oracle.dopri5_mlp(om, y0, t, rtol=1e-6, atol=1e-8)
    assert (st.n_attempts > st.n_accepted).any(), "the case must exercise rejections"
    assert np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_accepted == int(st.n_accepted.sum())
    # decreasing t_span (repair R5: s = -t)
    tr = t[::-1].copy()
    sol, _ = solve_fwd(px, torch, field, y0, tr, rtol=1e-6, atol=1e-8)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, tr, rtol=1e-6, atol=1e-8)
    assert np.array_equal(sol.cpu().numpy(), ref)
    # first_step / max_step / safety options (base_adaptive_solver_rk.py:32-49)
    kw = dict(rtol=1e-5, atol=1e-7, first_step=0.01, max_step=0.2, safety=0.8, ifactor=5.0, dfactor=0.3)
    sol, _ = solve_fwd(px, torch, field, y0, t, **kw)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, **kw)
    assert np.array_equal(sol.cpu().numpy(), ref)


def test_dopri5_ragged_batches(px, torch, oracle):
    """B = 1, a prime, one-over-a-warp: the per-CTA queue and tail handling."""
    field, om = both(px, oracle, spiral_weights(), "cube")
    t = cfg2_tspan(6)
    for B in (1, 31, 33, 1009):
        y0 = cfg2_y0(B, seed=B)
        sol, _ = solve_fwd(px, torch, field, y0, t)
        ref, _, _, _ = oracle.dopri5_mlp(om, y0, t)
        assert np.array_equal(sol.cpu().numpy(), ref), B


def test_dopri5_status_words(px, torch, oracle):
    field, _ = both(px, oracle, spiral_weights(), "cube")
    y0 = cfg2_y0(64)
    om = oracle.MLP(*spiral_weights(), pre="cube")
    t5 = np.array([0.0, 5.0], f32)
    with pytest.raises(AssertionError, match="max_num_steps"):
        solve_fwd(px, torch, field, y0, t5, max_num_steps=3)
    assert oracle.dopri5_mlp(om, y0, t5, max_num_steps=3)[3] == 3
    bad = y0.copy()
    bad[7, 0] = np.inf
    with pytest.raises(AssertionError):
        solve_fwd(px, torch, field, bad, cfg2_tspan(4))
    sol, s = solve_fwd(px, torch, field, bad, cfg2_tspan(4), check_status=False)
    ref, _, _, rc = oracle.dopri5_mlp(om, bad, cfg2_tspan(4))
    assert rc != 0 and s.read_stats().status == rc
    keep = np.arange(64) != 7  # the other trajectories are unaffected (one controller per trajectory)
    assert np.array_equal(sol.cpu().numpy()[:, keep], ref[:, keep])
    with pytest.raises(px.UnsupportedFieldError):  # no fused kernel for D=9: loud, no fallback
        f9, _ = both(px, oracle, fanin_weights(9, 8), "id")
        solve_fwd(px, torch, f9, np.zeros((4, 9), f32), cfg2_tspan(4))


# ------------------------------------------------------------------------------------------------
# dopri5 forward, large states (xde_tile_adaptive.cu): register-tiled field, one controller per trajectory
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,pre,B", [(64, 256, "id", 100), (64, 128, "cube", 77), (32, 256, "id", 65),
                                        (32, 128, "square", 130), (32, 64, "id", 64), (16, 64, "cube", 257)])
def test_dopri5_large_state_bit_exact_and_step_sequence(px, torch, oracle, d, h, pre, B):
    """Every tile geometry, ragged tiles (B not a multiple of TM), outputs by dense interpolation, identical
    accept / reject sequences (dt, ratio) per trajectory."""
    w = [1.5 * a for a in fanin_weights(d, h, seed=d + h)]
    field, om = both(px, oracle, w, pre)
    y0 = np.random.default_rng(d).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1.5, 6).astype(f32)
    kw = dict(rtol=1e-6, atol=1e-8)
    sol, s = solve_fwd(px, torch, field, y0, t, log_attempts=512, **kw)
    ref, st, _, rc = oracle.dopri5_mlp(om, y0, t, **kw)
    assert rc == 0
    assert np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum())
    assert s.stats.n_accepted == int(st.n_accepted.sum())
    assert s.stats.nfe == int(st.nfe.sum())
    rec, cnt = s.attempt_log.read()
    for b in (0, 1, B // 2, B - 1):
        _, _, lg, _ = oracle.dopri5_mlp(om, y0, t, log_traj=b, **kw)
        assert cnt[b] == len(lg)
        r = rec[b, :cnt[b]]
        assert np.array_equal(r.accepted, lg.accepted)
        assert np.array_equal(r.dt, lg.dt) and np.array_equal(r.t0, lg.t0)
        assert np.array_equal(r.ratio, lg.ratio)


def test_dopri5_large_state_rejections_reverse_options_status(px, torch, oracle):
    d, h = 32, 64
    w = [4.0 * a for a in fanin_weights(d, h, seed=9)]
    field, om = both(px, oracle, w, "id")
    y0 = np.random.default_rng(3).uniform(-1, 1, (97, d)).astype(f32)
    t = np.linspace(0, 2, 5).astype(f32)
    kw = dict(rtol=1e-5, atol=1e-7)
    sol, s = solve_fwd(px, torch, field, y0, t, **kw)
    ref, st, _, _ = oracle.dopri5_mlp(om, y0, t, **kw)
    assert (st.n_attempts > st.n_accepted).any(), "the case must exercise rejections"
    assert np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum()) and s.stats.n_accepted == int(st.n_accepted.sum())
    tr = t[::-1].copy()  # decreasing t_span (repair R5)
    sol, _ = solve_fwd(px, torch, field, y0, tr, **kw)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, tr, **kw)
    assert np.array_equal(sol.cpu().numpy(), ref)
    kw = dict(rtol=1e-5, atol=1e-7, first_step=0.01, max_step=0.2, safety=0.8, ifactor=5.0, dfactor=0.3)
    sol, _ = solve_fwd(px, torch, field, y0, t, **kw)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, **kw)
    assert np.array_equal(sol.cpu().numpy(), ref)
    # status words: max_num_steps, a non-finite trajectory leaves its tile mates untouched
    with pytest.raises(AssertionError, match="max_num_steps"):
        solve_fwd(px, torch, field, y0, t, max_num_steps=2, rtol=1e-7, atol=1e-9)
    bad = y0.copy()
    bad[5, 3] = np.inf
    sol, s = solve_fwd(px, torch, field, bad, t, check_status=False, rtol=1e-5, atol=1e-7)
    ref, _, _, rc = oracle.dopri5_mlp(om, bad, t, rtol=1e-5, atol=1e-7)
    assert rc != 0 and s.read_stats().status == rc
    keep = np.arange(97) != 5
    assert np.array_equal(sol.cpu().numpy()[:, keep], ref[:, keep])
    assert np.isnan(sol.cpu().numpy()[1:, 5]).all()


@pytest.mark.parametrize("name,rtol", [("Bosh3", 1e-5), ("Fehlberg2", 1e-4), ("AdaptiveHeun", 1e-4)])
@pytest.mark.parametrize("d,h,pre,B", [(64, 256, "id", 70), (32, 64, "cube", 130), (16, 64, "square", 33), (64, 128, "id", 5)])
def test_other_tableaux_large_state(px, torch, oracle, name, rtol, d, h, pre, B):
    """The tiled kernel with the tableau as data (stage count 3 / 2 / 1, FSAL or not, controller order 3 / 2):
    bit-exact incl. counters and a trajectory's attempt log; Dopri8 (13 stages) is loud about not fitting."""
    w = fanin_weights(d, h, seed=d + h)
    field, om = both(px, oracle, w, pre)
    y0 = np.random.default_rng(h).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 0.8, 4).astype(f32)
    method = {"Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "AdaptiveHeun": "adaptive_heun"}[name]
    xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), t)
    s = getattr(px, name)(xde=xde, y0=xde.y0, rtol=rtol, atol=rtol * 1e-2, log_attempts=1024)
    sol = s.integrate(t)
    ref, st, _, rc = oracle.adaptive_rk_mlp(method, om, y0, t, rtol=rtol, atol=rtol * 1e-2)
    assert rc == 0 and np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum()) and s.stats.n_accepted == int(st.n_accepted.sum())
    assert s.stats.nfe == int(st.nfe.sum())
    rec, cnt = s.attempt_log.read()
    b = B // 2
    _, _, lg, _ = oracle.adaptive_rk_mlp(method, om, y0, t, rtol=rtol, atol=rtol * 1e-2, log_traj=b)
    assert cnt[b] == len(lg) and rec[b, :cnt[b]].tobytes() == lg.tobytes()
    tr = t[::-1].copy()
    sol = getattr(px, name)(xde=px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), tr), y0=xde.y0, rtol=rtol,
                            atol=rtol * 1e-2).integrate(tr)
    ref, _, _, _ = oracle.adaptive_rk_mlp(method, om, y0, tr, rtol=rtol, atol=rtol * 1e-2)
    assert np.array_equal(sol.cpu().numpy(), ref)
    if name == "Bosh3" and d == 16:
        with pytest.raises(px.UnsupportedFieldError, match="stages"):
            px.Dopri8(xde=xde, y0=xde.y0, rtol=1e-6, atol=1e-8).integrate(t)
    # forced grid points on the tiled kernel (per-row indices; a jump re-evaluates f for the whole tile)
    step_t, jump_t = [0.21, 0.55, -1.0, 7.0], [0.4, 0.05]
    kw = dict(rtol=rtol, atol=rtol * 1e-2, step_t=step_t, jump_t=jump_t)
    s2 = getattr(px, name)(xde=xde, y0=xde.y0, log_attempts=1024, **kw)
    sol = s2.integrate(t)
    ref, st, _, rc = oracle.adaptive_rk_mlp(method, om, y0, t, **kw)
    assert rc == 0 and np.array_equal(sol.cpu().numpy(), ref)
    assert s2.stats.n_attempts == int(st.n_attempts.sum()) and s2.stats.nfe == int(st.nfe.sum())
    rec, cnt = s2.attempt_log.read()
    _, _, lg, _ = oracle.adaptive_rk_mlp(method, om, y0, t, log_traj=b, **kw)
    assert cnt[b] == len(lg) and rec[b, :cnt[b]].tobytes() == lg.tobytes()


@pytest.mark.parametrize("B", [1, 20, 1000, 40000])
def test_dopri5_batch_controller_reference_faithful(px, torch, oracle, B):
    """controller="batch": the reference's single global RMS norm and dt (utils/ode_utils.py:8-9).  The
    accept/reject sequence, every dt and error ratio, and the solution must equal the oracle's literal
    batch run.  (fp64 partial sums are added in a different -- fixed -- order than the oracle's
    sequential loop; a flipped fp32 rounding of the ratio is possible in principle and would show here.)"""
    field, om = both(px, oracle, spiral_weights(), "cube")
    y0 = cfg2_y0(B, seed=3)
    t = cfg2_tspan(32 if B <= 1000 else 8)
    sol, s = solve_fwd(px, torch, field, y0, t, controller="batch", log_attempts=512)
    ref, st, lg, rc = oracle.dopri5_mlp(om, y0, t, controller="batch")
    assert rc == 0
    rec, cnt = s.attempt_log.read()
    assert cnt[0] == len(lg)
    r = rec[0, :cnt[0]]
    assert np.array_equal(r.accepted, lg.accepted)
    assert np.array_equal(r.dt, lg.dt) and np.array_equal(r.t0, lg.t0)
    assert np.array_equal(r.ratio, lg.ratio)
    assert np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts[0]) * B and s.stats.nfe == int(st.nfe[0]) * B


def test_dopri5_batch_controller_rejections_reverse_and_status(px, torch, oracle):
    w = [3.0 * a for a in fanin_weights(2, 50, seed=5)]
    field, om = both(px, oracle, w, "id")
    y0 = np.random.default_rng(1).uniform(-1, 1, (777, 2)).astype(f32)
    for t in (np.linspace(0, 4, 9).astype(f32), np.linspace(4, 0, 9).astype(f32)):
        kw = dict(rtol=1e-6, atol=1e-8)
        sol, s = solve_fwd(px, torch, field, y0, t, controller="batch", **kw)
        ref, st, lg, rc = oracle.dopri5_mlp(om, y0, t, controller="batch", **kw)
        assert rc == 0 and (lg.accepted == 0).any(), "the case must exercise rejections"
        assert np.array_equal(sol.cpu().numpy(), ref)
        assert s.stats.n_accepted == int(st.n_accepted[0]) * 777
    with pytest.raises(AssertionError, match="max_num_steps"):
        solve_fwd(px, torch, field, y0, np.array([0.0, 50.0], f32), controller="batch", max_num_steps=2)
    bad = y0.copy()
    bad[5, 1] = np.nan
    with pytest.raises(AssertionError, match="non-finite"):  # one bad trajectory aborts the batch, as in the reference
        solve_fwd(px, torch, field, bad, np.linspace(0, 4, 9).astype(f32), controller="batch", first_step=0.01)


def test_dopri5_full_size_subset_parity(px, torch, oracle):
    """cfg2 at BASELINE size (B = 2^20): with one controller per trajectory every trajectory is
    independent of its neighbours, so a random subset of the full-size GPU result must equal the
    oracle run on that subset alone -- bit for bit."""
    field, om = both(px, oracle, spiral_weights(), "cube")
    B = 1 << 20
    y0, t = cfg2_y0(B), cfg2_tspan(10)
    sol, s = solve_fwd(px, torch, field, y0, t)
    idx = np.random.default_rng(7).choice(B, 4096, replace=False)
    ref, st, _, rc = oracle.dopri5_mlp(om, y0[idx], t)
    assert rc == 0
    got = sol[:, torch.from_numpy(idx).cuda()].cpu().numpy()
    assert np.array_equal(got, ref)
    assert s.stats.status == 0 and s.stats.n_accepted >= B and s.stats.nfe == 3 * B + 6 * s.stats.n_attempts
    assert np.array_equal(sol[0].cpu().numpy(), y0)


# ------------------------------------------------------------------------------------------------
# adjoint
# ------------------------------------------------------------------------------------------------
def loss_grad(sol_np):
    """loss = mean(|y_T|) (cfg2) -> grad_y [T,B,D]"""
    gy = np.zeros_like(sol_np)
    gy[-1] = np.sign(sol_np[-1]) / sol_np[-1].size
    return gy


@pytest.mark.parametrize("d,h,pre,B", [(2, 50, "cube", 300), (2, 20, "id", 64), (1, 40, "square", 100), (4, 64, "id", 50),
                                        (8, 32, "id", 70), (8, 64, "cube", 33), (3, 17, "id", 40)])
def test_adjoint_parity(px, torch, oracle, d, h, pre, B):
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = spiral_weights() if (d, h) == (2, 50) else fanin_weights(d, h, seed=h)
    field, om = both(px, oracle, w, pre)
    y0 = cfg2_y0(B) if d == 2 else np.random.default_rng(3).uniform(-1, 1, (B, d)).astype(f32)
    t = cfg2_tspan(6) if d == 2 else np.linspace(0, 1, 5).astype(f32)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t)
    gy = loss_grad(ref)
    gy[2] = 0.01 * np.random.default_rng(4).standard_normal(gy[2].shape).astype(f32)  # a mid-time cotangent
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=512)
    g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy)
    assert rc == 0
    s = stats.read()
    assert s.status == 0
    assert s.n_attempts == int(st_ref.n_attempts.sum()) and s.n_accepted == int(st_ref.n_accepted.sum())
    assert s.nfe == int(st_ref.nfe.sum())
    assert np.array_equal(a0.cpu().numpy(), a_ref), "adjoint state dL/dy0 must be bit-exact"
    scale = np.abs(g_ref).max()
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * scale)
    rec, cnt = log.read()
    for b in (0, B // 2, B - 1):
        _, _, _, lg, _ = oracle.dopri5_mlp_adjoint(om, t, ref, gy, log_traj=b)
        assert cnt[b] == len(lg)
        r = rec[b, :cnt[b]]
        assert np.array_equal(r.accepted, lg.accepted) and np.array_equal(r.dt, lg.dt)
        assert np.array_equal(r.ratio, lg.ratio)


def test_adjoint_with_rejections_and_replay(px, torch, oracle):
    """A stiffer field so that the ADJOINT controller rejects steps: the kernel folds parameter-gradient
    contributions eagerly and takes a rejected attempt back with a REPLAY block -- the state path must
    stay bit-exact and the gradients within rtol 1e-5."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = [3.0 * a for a in fanin_weights(2, 50, seed=5)]
    field, om = both(px, oracle, w, "id")
    B = 500
    y0 = np.random.default_rng(1).uniform(-1, 1, (B, 2)).astype(f32)
    t = np.linspace(0, 4, 5).astype(f32)
    kw = dict(rtol=1e-6, atol=1e-8)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, **kw)
    gy = 0.01 * np.random.default_rng(4).standard_normal(ref.shape).astype(f32)
    g, a0, stats, _ = adjoint_backward(field, t, ref, gy, return_adj_y0=True, **kw)
    g_ref, a_ref, st_ref, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, **kw)
    assert rc == 0
    n_rej = int((st_ref.n_attempts - st_ref.n_accepted).sum())
    assert n_rej > B // 10, f"the case must exercise rejections (got {n_rej})"
    s = stats.read()
    assert (s.n_attempts, s.n_accepted) == (int(st_ref.n_attempts.sum()), int(st_ref.n_accepted.sum()))
    assert np.array_equal(a0.cpu().numpy(), a_ref)
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=2e-6 * np.abs(g_ref).max())


@pytest.mark.parametrize("norm", ["seminorm", "mixed"])
@pytest.mark.parametrize("B,d,h,pre", [(1, 2, 50, "cube"), (300, 2, 50, "cube"), (5000, 2, 50, "cube"), (130, 4, 32, "id"),
                                       (77, 1, 40, "square")])
def test_adjoint_batch_controller_reference_default(px, torch, oracle, norm, B, d, h, pre):
    """controller="batch": one dt / one error norm for the whole augmented state -- with norm="mixed" this is
    the reference's DEFAULT odeint_adjoint configuration (functional/odeint_adjoint.py:284-291).
    At rtol 1e-7 (below fp32 epsilon) the parameter-gradient error estimate of the mixed norm is the rounding noise
    of the batch sum (ratio ~1e-2), so the accept/reject sequence depends on how that sum is taken.
    Round 2 specifies it order-independently (32-trajectory fp32 fma chains added exactly in 128-bit fixed point,
    one rounding: oracle adj_rhs, csrc/xde_fixed128.cuh): the WHOLE (dt, ratio, accept) sequence, dL/dy0 and the
    parameter gradients are now bit-identical to the oracle's for both norms."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = spiral_weights() if (d, h) == (2, 50) else fanin_weights(d, h, seed=h)
    field, om = both(px, oracle, w, pre)
    y0 = cfg2_y0(B, seed=4) if d == 2 else np.random.default_rng(3).uniform(-1, 1, (B, d)).astype(f32)
    t = cfg2_tspan(6) if d == 2 else np.linspace(0, 1, 5).astype(f32)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, controller="batch")
    gy = loss_grad(ref)
    gy[2] = 0.01 * np.random.default_rng(4).standard_normal(gy[2].shape).astype(f32)
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=1024,
                                         controller="batch", adj_norm=norm)
    g_ref, a_ref, st_ref, lg, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, controller="batch", adj_norm=norm)
    assert rc == 0
    s = stats.read()
    assert s.status == 0
    rec, cnt = log.read()
    r = rec[0, :cnt[0]]
    assert cnt[0] == len(lg) and np.array_equal(r.accepted, lg.accepted)
    assert s.n_attempts == int(st_ref.n_attempts[0]) * B and s.nfe == int(st_ref.nfe[0]) * B
    assert np.array_equal(r.dt, lg.dt) and np.array_equal(r.ratio, lg.ratio)
    assert np.array_equal(a0.cpu().numpy(), a_ref)
    assert np.array_equal(g.cpu().numpy(), g_ref), "g_theta is a state of this solve: bit-exact like y and a"


def test_adjoint_batch_mixed_norm_with_rejections(px, torch, oracle):
    """The reference's default configuration at rtol 1e-4 / atol 1e-6 on a stiffer field, so that the mixed-norm
    controller REJECTS steps: the whole sequence including the rejections, and every result, bit for bit."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = [3.0 * a for a in fanin_weights(2, 50, seed=5)]
    field, om = both(px, oracle, w, "id")
    B = 400
    y0 = np.random.default_rng(1).uniform(-1, 1, (B, 2)).astype(f32)
    t = np.linspace(0, 4, 5).astype(f32)
    kw = dict(rtol=1e-4, atol=1e-6)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, controller="batch", **kw)
    gy = 0.01 * np.random.default_rng(4).standard_normal(ref.shape).astype(f32)
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=1024, controller="batch",
                                         adj_norm="mixed", **kw)
    g_ref, a_ref, st_ref, lg, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, controller="batch", adj_norm="mixed", **kw)
    assert rc == 0 and stats.read().status == 0
    rec, cnt = log.read()
    r = rec[0, :cnt[0]]
    assert (lg.accepted == 0).any(), "the case must exercise rejections"
    assert cnt[0] == len(lg) and np.array_equal(r.accepted, lg.accepted)
    assert np.array_equal(r.dt, lg.dt) and np.array_equal(r.ratio, lg.ratio)
    assert np.array_equal(a0.cpu().numpy(), a_ref) and np.array_equal(g.cpu().numpy(), g_ref)


def test_adjoint_batch_mixed_norm_at_baseline_size_subset(px, torch, oracle):
    """B = 2^16 rows of the cfg2 batch under the reference-default controller: 2048 chain blocks, every CTA of the
    cooperative grid contributes to the 128-bit sums -- still the oracle's sequence and gradients bit for bit."""
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    field, om = both(px, oracle, spiral_weights(), "cube")
    B = 1 << 16
    y0, t = cfg2_y0(B), cfg2_tspan(4)
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, controller="batch")
    gy = loss_grad(ref)
    g, a0, stats, log = adjoint_backward(field, t, ref, gy, return_adj_y0=True, log_attempts=256, controller="batch",
                                         adj_norm="mixed")
    g_ref, a_ref, st_ref, lg, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, controller="batch", adj_norm="mixed")
    assert rc == 0 and stats.read().status == 0
    rec, cnt = log.read()
    r = rec[0, :cnt[0]]
    assert cnt[0] == len(lg) and np.array_equal(r.dt, lg.dt) and np.array_equal(r.ratio, lg.ratio)
    assert np.array_equal(a0.cpu().numpy(), a_ref) and np.array_equal(g.cpu().numpy(), g_ref)


def test_odeint_adjoint_reference_defaults_via_options(px, torch, oracle):
    """options={"controller": "batch"}: forward with the global controller, backward with the reference's default
    mixed adjoint norm -- the literal reference configuration, end to end through the public API."""
    w = spiral_weights()
    tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in w]
    field = px.MLPField(*tw, pre="cube")
    om = oracle.MLP(*w, pre="cube")
    y0 = cfg2_y0(256, seed=9)
    t = cfg2_tspan(5)
    sol = px.odeint_adjoint(field, torch.from_numpy(y0).cuda(), t, solver=px.Dopri5,
                            options={"norm": px.utils._rms_norm, "controller": "batch"})
    sol[-1].abs().mean().backward()
    ref, _, _, _ = oracle.dopri5_mlp(om, y0, t, controller="batch")
    assert np.array_equal(sol.detach().cpu().numpy(), ref)
    g_ref, _, _, _, _ = oracle.dopri5_mlp_adjoint(om, t, ref, loss_grad(ref), controller="batch", adj_norm="mixed")
    got = np.concatenate([p.grad.cpu().numpy().ravel() for p in tw])
    np.testing.assert_allclose(got, g_ref, rtol=1e-5, atol=2e-6 * np.abs(g_ref).max())
    assert px.odeint_adjoint.last["adj_norm"] == "mixed"


def test_adjoint_gradient_is_the_true_gradient(px, torch, oracle):
    """Independent of the oracle: finite differences of the GPU forward in fp32 are too noisy, so
    compare with the fp64 NumPy adjoint-free gradient (discretise-then-differentiate is not the same
    as the continuous adjoint; at rtol 1e-7 they agree to ~1e-4 relative)."""
    from oracle import oracle_np as onp
    from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

    w = spiral_weights()
    field, om = both(px, oracle, w, "cube")
    y0, t = cfg2_y0(32), cfg2_tspan(5)
    sol, _ = solve_fwd(px, torch, field, y0, t)
    sol_np = sol.cpu().numpy()
    gy = loss_grad(sol_np)
    g, _, _, _ = adjoint_backward(field, t, sol, gy)
    g = g.cpu().numpy()
    # fp64 central differences on a fixed-step fp64 RK4 of the same field
    def loss(params):
        fld = onp.MLPFieldNP(*params, "cube", np.float64)
        y = y0.astype(np.float64)
        tt = np.linspace(float(t[0]), float(t[-1]), 401)
        for i in range(400):
            hh = tt[i + 1] - tt[i]
            k1 = fld(tt[i], y)
            k2 = fld(tt[i] + hh / 2, y + hh / 2 * k1)
            k3 = fld(tt[i] + hh / 2, y + hh / 2 * k2)
            k4 = fld(tt[i + 1], y + hh * k3)
            y = y + hh / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        return np.abs(y).mean()

    params = [a.astype(np.float64) for a in w]
    flat_idx = [(0, (0, 3)), (0, (1, 17)), (1, (5,)), (2, (9, 1)), (2, (30, 0)), (3, (1,))]
    off = [0, 100, 150, 250]
    for pi, ix in flat_idx:
        e = 1e-5
        pp = [a.copy() for a in params]
        pp[pi][ix] += e
        pm = [a.copy() for a in params]
        pm[pi][ix] -= e
        fd = (loss(pp) - loss(pm)) / (2 * e)
        k = off[pi] + int(np.ravel_multi_index(ix, params[pi].shape))
        assert abs(g[k] - fd) <= 2e-3 * abs(fd) + 1e-5 * np.abs(g).max(), (pi, ix, g[k], fd)


def test_odeint_adjoint_autograd_surface(px, torch, oracle):
    """odeint_adjoint(...).backward(): grads for the parameters only, y0 gets None
    (functional/odeint_adjoint.py:167)."""
    w = spiral_weights()
    tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in w]
    field = px.MLPField(*tw, pre="cube")
    om = oracle.MLP(*w, pre="cube")
    y0 = torch.from_numpy(cfg2_y0(128)).cuda().requires_grad_(True)
    t = cfg2_tspan(5)
    sol = px.odeint_adjoint(field, y0, t, solver=px.Dopri5)
    assert tuple(sol.shape) == (5, 128, 2)
    loss = sol[-1].abs().mean()
    loss.backward()
    assert y0.grad is None
    ref, _, _, _ = oracle.dopri5_mlp(om, y0.detach().cpu().numpy(), t)
    assert np.array_equal(sol.detach().cpu().numpy(), ref)
    g_ref, _, _, _, _ = oracle.dopri5_mlp_adjoint(om, t, ref, loss_grad(ref))
    got = np.concatenate([p.grad.cpu().numpy().ravel() for p in tw])
    np.testing.assert_allclose(got, g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max())


def test_odeint_adjoint_deferred_status_check(px, torch, oracle):
    """options={"check_status": "deferred"}: no host round trip between the two solves; same numbers, and the
    forward assertion (here max_num_steps) surfaces in backward() instead of in the call."""
    w = spiral_weights()
    y0, t = torch.from_numpy(cfg2_y0(500)).cuda(), cfg2_tspan(6)

    def grads(**kw):
        tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in w]
        sol = px.odeint_adjoint(px.MLPField(*tw, pre="cube"), y0, t, solver=px.Dopri5, **kw)
        sol[-1].abs().mean().backward()
        return sol.detach(), torch.cat([p.grad.reshape(-1) for p in tw])

    sol_a, g_a = grads(options={"check_status": True})
    sol_b, g_b = grads(options={"check_status": "deferred"})
    sol_c, g_c = grads()  # the default of a training step (parameters require grad) is the deferred check
    assert torch.equal(sol_a, sol_c) and torch.equal(g_b, g_c)
    assert torch.equal(sol_a, sol_b)
    np.testing.assert_allclose(g_b.cpu().numpy(), g_a.cpu().numpy(), rtol=1e-6, atol=1e-7 * float(g_a.abs().max()))
    tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in w]
    field = px.MLPField(*tw, pre="cube")
    with pytest.raises(AssertionError, match="max_num_steps"):
        px.odeint_adjoint(field, y0, np.array([0.0, 5.0], f32), solver=px.Dopri5,
                          options={"max_num_steps": 3, "check_status": True})
    with pytest.raises(AssertionError, match="max_num_steps"):  # nothing requires grad: checked by the call
        px.odeint_adjoint(px.MLPField(*w, pre="cube"), y0, np.array([0.0, 5.0], f32), solver=px.Dopri5,
                          options={"max_num_steps": 3})
    sol = px.odeint_adjoint(field, y0, np.array([0.0, 5.0], f32), solver=px.Dopri5,
                            options={"max_num_steps": 3, "check_status": "deferred"})  # does not raise here ...
    with pytest.raises(AssertionError, match="max_num_steps"):
        sol[-1].nan_to_num().sum().backward()                                        # ... but here


# ------------------------------------------------------------------------------------------------
# fixed-grid solvers, SDE
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["Euler", "RK4"])
@pytest.mark.parametrize("d,h,pre", [(2, 50, "cube"), (4, 32, "id"), (8, 48, "square"), (5, 20, "id"), (7, 33, "cube")])
def test_fixed_solvers_bit_exact(px, torch, oracle, solver, d, h, pre):
    w = spiral_weights() if d == 2 else fanin_weights(d, h)
    field, om = both(px, oracle, w, pre)
    B = 515
    y0 = cfg2_y0(B) if d == 2 else np.random.default_rng(0).uniform(-1, 1, (B, d)).astype(f32)
    t = cfg2_tspan(32) if d == 2 else np.linspace(0, 1, 21).astype(f32)
    y0_3d = torch.from_numpy(y0).cuda().reshape(B, 1, d)  # [B,1,D] -> [B,T,D] (base_fixed_solver.py:143)
    sol = px.odeint(field, y0_3d, t, getattr(px, solver))
    ref = oracle.fixed_mlp(solver.lower(), om, y0, t)
    assert tuple(sol.shape) == (B, t.size, d)
    assert np.array_equal(sol.cpu().numpy(), ref)
    # y0 [B,D] -> concat(axis=-2) gives [T*B, D]
    sol2 = px.odeint(field, torch.from_numpy(y0).cuda(), t, getattr(px, solver))
    assert np.array_equal(sol2.cpu().numpy().reshape(t.size, B, d), ref.transpose(1, 0, 2))


@pytest.mark.parametrize("scheme", ["em", "milstein"])
@pytest.mark.parametrize("d", [4, 3, 7])
def test_sde_supplied_increments_bit_exact(px, torch, oracle, scheme, d):
    h, B = 32, 700
    wf, wg = fanin_weights(d, h, seed=2), fanin_weights(d, h, seed=3)
    f, of = both(px, oracle, wf, "cube")
    g, og = both(px, oracle, wg, "square")
    rng = np.random.default_rng(2)
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 17).astype(f32)
    dW = (np.sqrt(1 / 16) * rng.standard_normal((16, B, d))).astype(f32)
    sol = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler,
                    options={"bm_increments": torch.from_numpy(dW).cuda(), "scheme": scheme})
    ref = oracle.sde_mlp(scheme, of, og, y0, t, dW)
    assert np.array_equal(sol.cpu().numpy(), ref)
    with pytest.raises(ValueError):  # increments are mandatory: no host Brownian tree on this path
        px.sdeint(f, g, torch.from_numpy(y0).cuda(), t, px.Euler)


# ------------------------------------------------------------------------------------------------
# the other embedded Runge-Kutta pairs: table-driven kernel, csrc/xde_adaptive_rk.cu
# ------------------------------------------------------------------------------------------------
def test_table_driven_kernel_equals_tuned_dopri5(px, torch, oracle):
    """The generic kernel handed the Dormand-Prince tableau == the tuned Dopri5 kernel, bit for bit:
    states, counters and every record of the attempt logs."""
    from paddlexde_b200.solver.adaptive_solver import _Dopri5Table

    field, _ = both(px, oracle, spiral_weights(), "cube")
    y0 = cfg2_y0(777)
    t = cfg2_tspan(12)
    res = []
    for cls in (px.Dopri5, _Dopri5Table):
        xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), t)
        s = cls(xde=xde, y0=xde.y0, rtol=1e-7, atol=1e-9, log_attempts=64)
        sol = s.integrate(t).cpu().numpy()
        rec, cnt = s.attempt_log.read()
        res.append((sol, s.stats, rec, cnt))
    (a, sa, ra, ca), (b, sb, rb, cb) = res
    assert np.array_equal(a, b) and sa == sb and np.array_equal(ca, cb)
    for i in range(0, 777, 37):
        assert ra[i, :ca[i]].tobytes() == rb[i, :cb[i]].tobytes()


@pytest.mark.parametrize("name,rtol", [("Bosh3", 1e-6), ("Fehlberg2", 1e-4), ("AdaptiveHeun", 1e-4), ("Dopri8", 1e-7),
                                       ("Dopri8", 1e-5)])
@pytest.mark.parametrize("d,h,pre,B", [(2, 50, "cube", 333), (4, 32, "id", 65), (1, 16, "square", 31), (6, 24, "id", 40)])
def test_other_tableaux_bit_exact_and_step_sequence(px, torch, oracle, name, rtol, d, h, pre, B):
    w = spiral_weights() if d == 2 else fanin_weights(d, h, seed=d)
    field, om = both(px, oracle, w, pre)
    rng = np.random.default_rng(d + B)
    y0 = (cfg2_y0(B) if d == 2 else rng.uniform(-1, 1, (B, d))).astype(f32)
    t = np.linspace(0, 1.0, 6).astype(f32)
    method = {"Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "AdaptiveHeun": "adaptive_heun", "Dopri8": "dopri8"}[name]
    xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), t)
    s = getattr(px, name)(xde=xde, y0=xde.y0, rtol=rtol, atol=rtol * 1e-2, log_attempts=4096)
    sol = s.integrate(t)
    ref, st, _, rc = oracle.adaptive_rk_mlp(method, om, y0, t, rtol=rtol, atol=rtol * 1e-2)
    assert rc == 0
    assert np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum()) and s.stats.n_accepted == int(st.n_accepted.sum())
    assert s.stats.nfe == int(st.nfe.sum())
    rec, cnt = s.attempt_log.read()
    for b in (0, B // 2, B - 1):  # identical accept/reject sequence, dt and error ratio
        _, _, lg, _ = oracle.adaptive_rk_mlp(method, om, y0, t, log_traj=b, rtol=rtol, atol=rtol * 1e-2)
        assert cnt[b] == len(lg) <= 4096
        r = rec[b, :cnt[b]]
        assert np.array_equal(r.accepted, lg.accepted) and np.array_equal(r.dt, lg.dt)
        assert np.array_equal(r.ratio, lg.ratio)


def test_other_tableaux_reverse_time_first_step_and_status(px, torch, oracle):
    field, om = both(px, oracle, spiral_weights(), "cube")
    y0 = cfg2_y0(50)
    t = np.linspace(0.5, 0.0, 5).astype(f32)  # decreasing: repair R5
    sol = px.odeint(field, torch.from_numpy(y0).cuda(), t, px.Bosh3, rtol=1e-5, atol=1e-7,
                    options={"first_step": 0.01})
    ref, _, _, rc = oracle.adaptive_rk_mlp("bosh3", om, y0, t, rtol=1e-5, atol=1e-7, first_step=0.01)
    assert rc == 0 and np.array_equal(sol.cpu().numpy(), ref)
    with pytest.raises(AssertionError, match="max_num_steps"):
        px.odeint(field, torch.from_numpy(y0).cuda(), np.linspace(0, 5, 3).astype(f32), px.AdaptiveHeun,
                  rtol=1e-6, atol=1e-8, options={"max_num_steps": 5})
    with pytest.raises(px.UnsupportedFieldError):  # reference-faithful global controller: Dopri5 only
        px.odeint(field, torch.from_numpy(y0).cuda(), np.linspace(0, 1, 3).astype(f32), px.Bosh3,
                  options={"controller": "batch"})


@pytest.mark.parametrize("name", ["Dopri5", "Bosh3", "Fehlberg2", "Dopri8"])
@pytest.mark.parametrize("reverse", [False, True])
def test_step_t_and_jump_t_forced_grid_points(px, torch, oracle, name, reverse):
    """solver kwargs step_t / jump_t (base_adaptive_solver_rk.py:94-114, 209-224, 263-273): attempts are
    clamped to the next forced point, a jump re-evaluates f; bit-exact incl. the attempt log and nfe."""
    field, om = both(px, oracle, spiral_weights(), "cube")
    y0 = cfg2_y0(200)
    t = (np.linspace(1.0, 0.0, 4) if reverse else np.linspace(0, 1.5, 5)).astype(f32)
    step_t, jump_t = [0.33, 0.9, -1.0, 0.05, 7.0], [0.5, 1.2, 0.051]
    method = {"Dopri5": "dopri5", "Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "Dopri8": "dopri8"}[name]
    xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda(), t)
    s = getattr(px, name)(xde=xde, y0=xde.y0, rtol=1e-5, atol=1e-7, step_t=torch.tensor(step_t), jump_t=jump_t,
                          log_attempts=512)
    sol = s.integrate(t)
    ref, st, _, rc = oracle.adaptive_rk_mlp(method, om, y0, t, rtol=1e-5, atol=1e-7, step_t=step_t, jump_t=jump_t)
    plain, stp, _, _ = oracle.adaptive_rk_mlp(method, om, y0, t, rtol=1e-5, atol=1e-7)
    assert rc == 0 and np.array_equal(sol.cpu().numpy(), ref)
    assert s.stats.n_attempts == int(st.n_attempts.sum()) > int(stp.n_attempts.sum())  # the points cost steps
    assert s.stats.nfe == int(st.nfe.sum())
    rec, cnt = s.attempt_log.read()
    _, _, lg, _ = oracle.adaptive_rk_mlp(method, om, y0, t, log_traj=17, rtol=1e-5, atol=1e-7, step_t=step_t,
                                         jump_t=jump_t)
    r = rec[17, :cnt[17]]
    assert cnt[17] == len(lg) and np.array_equal(r.dt, lg.dt) and np.array_equal(r.t0, lg.t0)
    assert np.array_equal(r.accepted, lg.accepted)
    ends = np.round((r.t0 + r.dt)[r.accepted == 1], 5)   # an accepted attempt ends on every forced point in range
    inside = [v for v in step_t + jump_t if min(t[0], t[-1]) < v < max(t[0], t[-1])]
    assert all(np.any(np.isclose(ends, v, atol=2e-6)) for v in inside)


def test_fixed_interp_cubic_is_the_same_kernel(px, torch, oracle):
    field, om = both(px, oracle, spiral_weights(), "cube")
    y0 = cfg2_y0(64)
    t = np.linspace(0, 1, 9).astype(f32)
    yd = torch.from_numpy(y0).cuda().reshape(64, 1, 2)
    a = px.odeint(field, yd, t, px.RK4, options={"interp": "cubic"})
    assert np.array_equal(a.cpu().numpy(), oracle.fixed_mlp("rk4", om, y0, t))
    with pytest.raises(NotImplementedError):  # cubic interpolation on a constructed grid needs f at the grid points
        px.odeint(field, yd, t, px.RK4, options={"step_size": 0.01, "interp": "cubic"})


@pytest.mark.parametrize("d,h,pre,B", [(2, 50, "cube", 100), (8, 48, "square", 33), (64, 256, "id", 70), (32, 64, "cube", 129)])
def test_midpoint_fixed_solver(px, torch, oracle, d, h, pre, B):
    """fixed_solver/midpoint.py:7-18 on all three fixed-grid kernels: bit-exact on the FP32 ones
    (small-state and register-tiled), rtol 1e-5 on the tensor-core one."""
    field, om = both(px, oracle, fanin_weights(d, h, seed=d + 1), pre)
    y0 = np.random.default_rng(d).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 9).astype(f32)
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    ref = oracle.fixed_mlp("midpoint", om, y0, t)
    sol = px.odeint(field, yd, t, px.Midpoint, options={"math": "fp32"})
    assert np.array_equal(sol.cpu().numpy(), ref)
    if d >= 16:
        tens = px.odeint(field, yd, t, px.Midpoint, options={"math": "tensor"})
        assert _close(tens.cpu().numpy(), ref, rtol=1e-5)


# ------------------------------------------------------------------------------------------------
# device-side Brownian increments (Philox4x32-10 + Box-Muller)
# ------------------------------------------------------------------------------------------------
def test_brownian_increments_generator(px, torch, oracle):
    from oracle.philox_np import brownian_increments as ref_inc
    from paddlexde_b200.utils.brownian import brownian_increments

    t = np.array([0.0, 0.1, 0.35, 1.0], f32)
    for D in (1, 2, 6, 32):
        dev = brownian_increments(11, t, 500, D).cpu().numpy()
        ref = ref_inc(11, t, 500, D)
        assert dev.shape == ref.shape == (3, 500, D)
        np.testing.assert_allclose(dev, ref, rtol=1e-3, atol=2e-5)  # same uint32 stream; MUFU vs libm log/sin/cos
    whole = brownian_increments(5, t, 1000, 8)
    part = brownian_increments(5, t, 300, 8, offset=700)
    assert torch.equal(part, whole[:, 700:])                         # shard == slice of the whole batch
    big = brownian_increments(1, np.array([0.0, 0.25], f32), 1 << 20, 8)[0]
    assert abs(float(big.mean())) < 1e-3 and abs(float(big.var()) - 0.25) < 1e-3
    c = float((big[:, 0] * big[:, 1]).mean())
    assert abs(c) < 1e-3                                             # components are independent


@pytest.mark.parametrize("d,h,B,math,scheme", [(4, 32, 700, "fp32", "em"), (4, 32, 300, "fp32", "milstein"),
                                               (32, 64, 1000, "fp32", "em"), (32, 64, 1000, "tensor", "em"),
                                               (32, 64, 148 * 256 + 5, "tensor", "em"), (16, 64, 130, "tensor", "em")])
def test_sdeint_with_generated_increments(px, torch, oracle, d, h, B, math, scheme):
    """options={"bm_seed": s}: increments generated inside the kernel == the same increments written to a
    table and supplied (bit for bit, every kernel family); the FP32 kernels then equal the oracle run on that
    table; a sharded run with bm_offset equals the unsharded one."""
    from paddlexde_b200.utils.brownian import brownian_increments

    f, of = both(px, oracle, fanin_weights(d, h, seed=2), "cube")
    g, og = both(px, oracle, fanin_weights(d, h, seed=3), "square")
    y0 = np.random.default_rng(4).uniform(-1, 1, (B, d)).astype(f32)
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    t = np.linspace(0, 1, 9).astype(f32)
    opt = {"math": math, "scheme": scheme}
    gen = px.sdeint(f, g, yd, t, px.Euler, options={"bm_seed": 2024, **opt})
    table = brownian_increments(2024, t, B, d)
    sup = px.sdeint(f, g, yd, t, px.Euler, options={"bm_increments": table, **opt})
    assert torch.equal(gen, sup)
    if math == "fp32" and B <= 2000:
        ref = oracle.sde_mlp(scheme, of, og, y0, t, table.cpu().numpy())
        assert np.array_equal(gen.cpu().numpy(), ref)
    h0 = B // 3
    parts = [px.sdeint(f, g, yd[a:b], t, px.Euler, options={"bm_seed": 2024, "bm_offset": a, **opt})
             for a, b in ((0, h0), (h0, B))]
    assert torch.equal(torch.cat(parts), gen)
    with pytest.raises(ValueError):
        px.sdeint(f, g, yd, t, px.Euler, options={"bm_seed": 1, "bm_increments": table})


# ------------------------------------------------------------------------------------------------
# sdeint_adjoint: the exact adjoint of the Euler-Maruyama recursion (parity unpinned:
# the reference's own backward is a placeholder -- checked against the oracle and against fp64 autograd)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,B,T", [(2, 50, 333, 9), (4, 32, 65, 17), (1, 16, 31, 5), (8, 64, 100, 6), (4, 21, 40, 4),
                                     (2, 96, 7, 3)])
def test_sde_adjoint_matches_oracle(px, torch, oracle, d, h, B, T):
    from paddlexde_b200.functional.sdeint_adjoint import sde_adjoint_backward
    from paddlexde_b200.utils.brownian import brownian_increments

    f, of = both(px, oracle, fanin_weights(d, h, seed=2), "cube")
    g, og = both(px, oracle, fanin_weights(d, h, seed=3), "square")
    rng = np.random.default_rng(d * 100 + h)
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, T).astype(f32)
    table = brownian_increments(77, t, B, d)
    dW = table.cpu().numpy()
    sol = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler, options={"bm_increments": table})
    ref = oracle.sde_mlp("em", of, og, y0, t, dW)
    assert np.array_equal(sol.cpu().numpy(), ref)
    gy = (rng.standard_normal(ref.shape) / ref.size).astype(f32)
    gf_r, gg_r, a0_r = oracle.sde_mlp_adjoint(of, og, t, ref, gy, dW)
    gyd = torch.from_numpy(gy).cuda()
    gf, gg, a0 = sde_adjoint_backward(f, g, t, sol, gyd, bm_increments=table, return_adj_y0=True)
    assert np.array_equal(a0.cpu().numpy(), a0_r)                     # adjoint state: bit for bit
    for got, want in ((gf, gf_r), (gg, gg_r)):                          # parameter gradients: summation order
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
    # increments regenerated inside the backward kernel == the supplied table
    gf2, gg2, a02 = sde_adjoint_backward(f, g, t, sol, gyd, bm_seed=77, return_adj_y0=True)
    assert torch.equal(a02, a0)
    np.testing.assert_allclose(gf2.cpu().numpy(), gf.cpu().numpy(), rtol=1e-6, atol=1e-7 * np.abs(gf_r).max())
    np.testing.assert_allclose(gg2.cpu().numpy(), gg.cpu().numpy(), rtol=1e-6, atol=1e-7 * np.abs(gg_r).max())


def test_sdeint_adjoint_autograd_surface_is_the_true_gradient(px, torch, oracle):
    """sdeint_adjoint(...).backward() against fp64 autograd through the same Euler-Maruyama recursion."""
    d, h, B, T = 2, 50, 257, 9
    wf, wg = fanin_weights(d, h, seed=2), fanin_weights(d, h, seed=3)
    rng = np.random.default_rng(5)
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, T).astype(f32)
    dW = (np.sqrt(1 / (T - 1)) * rng.standard_normal((T - 1, B, d))).astype(f32)
    tw = [torch.tensor(a, device="cuda", requires_grad=True) for a in (*wf, *wg)]
    f, g = px.MLPField(*tw[:4], pre="cube"), px.MLPField(*tw[4:], pre="square")
    sol = px.sdeint_adjoint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler,
                            options={"bm_increments": torch.from_numpy(dW).cuda()})
    assert tuple(sol.shape) == (B, T, d)
    (sol[:, -1].abs().mean() + 0.1 * (sol[:, T // 2] ** 2).mean()).backward()
    P = [torch.tensor(np.asarray(a, np.float64), requires_grad=True) for a in (*wf, *wg)]
    F = lambda y, w1, b1, w2, b2, p: torch.tanh((y ** p) @ w1 + b1) @ w2 + b2
    y = torch.tensor(y0.astype(np.float64))
    ys = [y]
    for n in range(T - 1):
        y = y + F(y, *P[:4], 3) * (float(t[n + 1]) - float(t[n])) + F(y, *P[4:], 2) * torch.tensor(dW[n].astype(np.float64))
        ys.append(y)
    (ys[-1].abs().mean() + 0.1 * (ys[T // 2] ** 2).mean()).backward()
    for a, b in zip(tw, P):
        ref = b.grad.numpy()
        np.testing.assert_allclose(a.grad.cpu().numpy(), ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max())
    with pytest.raises(ValueError):
        px.sdeint_adjoint(f, g, torch.from_numpy(y0).cuda(), t, px.Euler, options={"bm_seed": 1})


# ------------------------------------------------------------------------------------------------
# large states: the register-tiled FFMA2 kernels (cfg3: 64-256-64 RK4, cfg4: 32-64-32 Euler-Maruyama)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,B", [(64, 256, 100), (64, 128, 33), (64, 64, 130), (32, 256, 31), (32, 128, 65),
                                   (32, 64, 200), (16, 64, 77)])
@pytest.mark.parametrize("solver", ["Euler", "RK4"])
def test_tiled_fixed_solvers_bit_exact(px, torch, oracle, solver, d, h, B):
    field, om = both(px, oracle, fanin_weights(d, h, seed=d + h), "id" if d == 64 else "cube")
    y0 = np.random.default_rng(d).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 11).astype(f32)
    sol = px.odeint(field, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, getattr(px, solver),
                    options={"math": "fp32"})
    ref = oracle.fixed_mlp(solver.lower(), om, y0, t)
    assert tuple(sol.shape) == (B, t.size, d)
    assert np.array_equal(sol.cpu().numpy(), ref)


def test_tiled_cfg3_full_size_subset_and_stride(px, torch, oracle):
    """cfg3 at the per-GPU BASELINE size (B = 2^17, D = 64, MLP 64-256-64, RK4, 100 steps, every 10th
    point stored): trajectories are independent, so a random subset must equal the oracle bit for bit."""
    d, h, B = 64, 256, 1 << 17
    field, om = both(px, oracle, fanin_weights(d, h, seed=1), "id")
    y0 = np.random.default_rng(1).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 101).astype(f32)
    xde = px.xde.BaseODE(field, torch.from_numpy(y0).cuda().reshape(B, 1, d), t)
    sol = px.RK4(xde=xde, y0=xde.y0, rtol=1e-7, atol=1e-9, out_stride=10, math="fp32").integrate(t)
    assert tuple(sol.shape) == (B, 11, d)
    idx = np.random.default_rng(3).choice(B, 192, replace=False)
    ref = oracle.fixed_mlp("rk4", om, y0[idx], t)[:, ::10]
    assert np.array_equal(sol[torch.from_numpy(idx).cuda()].cpu().numpy(), ref)
    assert np.isfinite(sol.cpu().numpy()).all()


def test_tiled_sde_cfg4_shapes(px, torch, oracle):
    """cfg4: D = 32, two 32-64-32 nets (drift y**3, diffusion y**2), 16 steps, supplied increments."""
    d, h, B = 32, 64, 1000
    f, of = both(px, oracle, fanin_weights(d, h, seed=2), "cube")
    g, og = both(px, oracle, fanin_weights(d, h, seed=3), "square")
    rng = np.random.default_rng(2)
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 17).astype(f32)
    dW = (np.sqrt(1 / 16) * rng.standard_normal((16, B, d))).astype(f32)
    sol = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler,
                    options={"bm_increments": torch.from_numpy(dW).cuda(), "math": "fp32"})
    ref = oracle.sde_mlp("em", of, og, y0, t, dW)
    assert np.array_equal(sol.cpu().numpy(), ref)
    auto = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler,
                     options={"bm_increments": torch.from_numpy(dW).cuda()})  # default: tensor cores
    assert _close(auto.cpu().numpy(), ref, rtol=1e-5) and not np.array_equal(auto.cpu().numpy(), ref)
    # Milstein at this shape (round 2): the FP32 tiles, bit-exact (tests/test_gpu_round2.py covers more shapes)
    mil = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler,
                    options={"bm_increments": torch.from_numpy(dW).cuda(), "scheme": "milstein"})
    assert np.array_equal(mil.cpu().numpy(), oracle.sde_mlp("milstein", of, og, y0, t, dW))
    with pytest.raises(px.UnsupportedFieldError):  # no kernel for this shape: loud, no fallback
        f48, _ = both(px, oracle, fanin_weights(48, 64), "id")
        px.odeint(f48, torch.zeros(4, 1, 48).cuda(), t, px.RK4)


# ------------------------------------------------------------------------------------------------
# large states on the tensor cores (csrc/xde_tc.cu): tcgen05 fp16-split 3-product GEMMs, fp32 accumulate.
# Not bit-exact by construction; the north-star tolerance (rtol 1e-5, with an atol scaled to the state:
# SURVEY appendix C take-away 3) is written in every assert.
# ------------------------------------------------------------------------------------------------
def _close(a, b, rtol=1e-5):
    scale = float(np.abs(b).max())
    return np.allclose(a, b, rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("d,h,B", [(64, 256, 100), (64, 128, 33), (64, 64, 130), (32, 256, 31), (32, 128, 257),
                                   (32, 64, 200), (16, 64, 77), (64, 256, 1)])
@pytest.mark.parametrize("solver", ["Euler", "RK4"])
def test_tensor_fixed_solvers_match_oracle(px, torch, oracle, solver, d, h, B):
    field, om = both(px, oracle, fanin_weights(d, h, seed=d + h), "id" if d == 64 else "cube")
    y0 = np.random.default_rng(d).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 11).astype(f32)
    sol = px.odeint(field, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, getattr(px, solver),
                    options={"math": "tensor"})
    ref = oracle.fixed_mlp(solver.lower(), om, y0, t)
    assert tuple(sol.shape) == (B, t.size, d)
    got = sol.cpu().numpy()
    assert np.array_equal(got[:, 0], y0)  # the initial row is copied, not computed
    assert _close(got, ref, rtol=1e-5)


def test_tensor_path_is_as_accurate_as_fp32(px, torch, oracle):
    """Against an fp64 evaluation of the same scheme the tensor path's error must be of the size of the
    FP32 path's own rounding error (measured ratio 0.6-1.9; bound 4)."""
    d, h, B = 64, 256, 512
    w = fanin_weights(d, h, seed=7)
    field = px.MLPField(*w, pre="id")
    y0 = np.random.default_rng(7).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 11).astype(f32)
    w1, b1, w2, b2 = [np.asarray(a, dtype=np.float64) for a in w]
    f = lambda y: np.tanh(y @ w1 + b1) @ w2 + b2
    y = y0.astype(np.float64)
    for i in range(1, t.size):  # rk4_alt_step_func, base_fixed_solver.py:166-197, in fp64
        dt = float(t[i]) - float(t[i - 1])
        k1 = f(y); k2 = f(y + dt / 3 * k1); k3 = f(y + dt * (k1 - k2 / 3)); k4 = f(y + dt * (k1 - k2 + k3))
        y = y + dt * (k1 + 3 * k2 + 3 * k3 + k4) / 8
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    e = {}
    for math in ("tensor", "fp32"):
        sol = px.odeint(field, yd, t, px.RK4, options={"math": math}).cpu().numpy()[:, -1]
        e[math] = np.abs(sol - y).max()
    assert e["tensor"] <= 4 * e["fp32"] + 1e-7, e


def test_tensor_cfg3_full_size_subset_and_stride(px, torch, oracle):
    """cfg3 at the per-GPU BASELINE size (B = 2^17, 64-256-64, RK4, 100 steps, every 10th point stored)
    on the tensor path: a random subset against the oracle at rtol 1e-5, and against the FP32 kernel."""
    d, h, B = 64, 256, 1 << 17
    field, om = both(px, oracle, fanin_weights(d, h, seed=1), "id")
    y0 = np.random.default_rng(1).uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 101).astype(f32)
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    sol = px.odeint(field, yd, t, px.RK4, options={"math": "tensor", "out_stride": 10})
    assert tuple(sol.shape) == (B, 11, d)
    idx = np.random.default_rng(3).choice(B, 192, replace=False)
    ref = oracle.fixed_mlp("rk4", om, y0[idx], t)[:, ::10]
    assert _close(sol[torch.from_numpy(idx).cuda()].cpu().numpy(), ref, rtol=1e-5)
    exact = px.odeint(field, yd, t, px.RK4, options={"math": "fp32", "out_stride": 10})
    scale = float(exact.abs().max())
    assert float((sol - exact).abs().max()) <= 1e-5 * scale
    assert bool(torch.isfinite(sol).all())


@pytest.mark.parametrize("d,h,B", [(32, 64, 1000), (32, 128, 130), (64, 64, 129), (16, 64, 5)])
def test_tensor_sde_matches_oracle(px, torch, oracle, d, h, B):
    """cfg4 shapes on the tensor path: two nets (drift y**3, diffusion y**2), 16 steps, supplied increments."""
    f, of = both(px, oracle, fanin_weights(d, h, seed=2), "cube")
    g, og = both(px, oracle, fanin_weights(d, h, seed=3), "square")
    rng = np.random.default_rng(2)
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    t = np.linspace(0, 1, 17).astype(f32)
    dW = (np.sqrt(1 / 16) * rng.standard_normal((16, B, d))).astype(f32)
    sol = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler,
                    options={"bm_increments": torch.from_numpy(dW).cuda(), "math": "tensor"})
    ref = oracle.sde_mlp("em", of, og, y0, t, dW)
    assert _close(sol.cpu().numpy(), ref, rtol=1e-5)


@pytest.mark.parametrize("kind,d,h", [("RK4", 32, 64), ("sde", 32, 64), ("Midpoint", 16, 64), ("Euler", 32, 128),
                                      ("sde", 16, 64), ("RK4", 64, 64)])
def test_tensor_many_tiles_per_cta(px, torch, oracle, kind, d, h):
    """More tile pairs than SMs: small fields (<= 256 TMEM columns, D <= 32) run fixed_tc2_kernel, two tiles
    in flight per CTA, the last pair half empty and the last tile ragged; (64, 64) stays on the one-tile
    kernel with several tiles per CTA.  Whole batch against the FP32 kernel, a subset against the oracle."""
    B = 148 * 128 * 2 + 77
    rng = np.random.default_rng(11)
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    idx = rng.choice(B, 160, replace=False)
    t = np.linspace(0, 1, 6).astype(f32)
    if kind != "sde":
        field, om = both(px, oracle, fanin_weights(d, h, seed=5), "cube")
        S = getattr(px, kind)
        sol = px.odeint(field, yd, t, S, options={"math": "tensor"})
        exact = px.odeint(field, yd, t, S, options={"math": "fp32"})
        ref = oracle.fixed_mlp(kind.lower(), om, y0[idx], t)
    else:
        f, of = both(px, oracle, fanin_weights(d, h, seed=2), "cube")
        g, og = both(px, oracle, fanin_weights(d, h, seed=3), "square")
        dW = (np.sqrt(1 / 5) * rng.standard_normal((5, B, d))).astype(f32)
        dWd = torch.from_numpy(dW).cuda()
        sol = px.sdeint(f, g, yd, t, px.Euler, options={"bm_increments": dWd, "math": "tensor"})
        exact = px.sdeint(f, g, yd, t, px.Euler, options={"bm_increments": dWd, "math": "fp32"})
        ref = oracle.sde_mlp("em", of, og, y0[idx], t, dW[:, idx])
    scale = float(exact.abs().max())
    assert float((sol - exact).abs().max()) <= 1e-5 * scale
    assert _close(sol[torch.from_numpy(idx).cuda()].cpu().numpy(), ref, rtol=1e-5)


def test_tensor_path_reports_inputs_outside_its_fp16_range(px, torch, oracle):
    """|pre(y)| >= 65504 (here y**3 with |y| ~ 50) cannot be split into fp16 operands: the kernels raise a
    status word; math="auto" then reruns on the FP32 kernels (== oracle bit for bit), math="tensor" raises."""
    d, h, B = 32, 64, 300
    field, om = both(px, oracle, fanin_weights(d, h, seed=9), "cube")
    y0 = np.random.default_rng(9).uniform(-1, 1, (B, d)).astype(f32)
    y0[17, 3] = 50.0
    t = np.linspace(0, 0.01, 3).astype(f32)
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    ref = oracle.fixed_mlp("rk4", om, y0, t)
    auto = px.odeint(field, yd, t, px.RK4)
    assert np.array_equal(auto.cpu().numpy(), ref)
    with pytest.raises(OverflowError):
        px.odeint(field, yd, t, px.RK4, options={"math": "tensor"})
    y0[17, 3] = 0.5                                    # back in range: the default is the tensor path again
    yd = torch.from_numpy(y0).cuda().reshape(B, 1, d)
    ok = px.odeint(field, yd, t, px.RK4)
    ref = oracle.fixed_mlp("rk4", om, y0, t)
    assert _close(ok.cpu().numpy(), ref, rtol=1e-5) and not np.array_equal(ok.cpu().numpy(), ref)


def test_tensor_path_is_loud_about_unsupported_shapes(px, torch, oracle):
    t = np.linspace(0, 1, 5).astype(f32)
    small, _ = both(px, oracle, fanin_weights(2, 50), "cube")
    with pytest.raises(px.UnsupportedFieldError):  # K = 2 is degenerate for an MMA: FP32 path only
        px.odeint(small, torch.zeros(4, 1, 2).cuda(), t, px.RK4, options={"math": "tensor"})
    with pytest.raises(ValueError):
        px.odeint(small, torch.zeros(4, 1, 2).cuda(), t, px.RK4, options={"math": "bf16"})


# ------------------------------------------------------------------------------------------------
# history gather / DDE
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["linear", "cubic", "bez"])
def test_history_gather_cfg5(px, torch, oracle, kind):
    """cfg5 shapes: his [8,307,288,3], 12 real-valued lags (+ edge queries on and off the grid)."""
    from paddlexde_b200.xde.base_dde import history_gather, history_gather_bwd

    rng = np.random.default_rng(5)
    his = rng.uniform(-1, 1, (8, 307, 288, 3)).astype(f32)
    his[..., 1:] = np.round(his[..., 1:] * 10)
    span = np.arange(288, dtype=f32)
    lags = (np.arange(12) + rng.uniform(0, 1, 12)).astype(f32)
    for lg in (lags, np.array([0.0, 1.0, 286.0, 286.5, 287.0, 3.25], f32), np.full(12, 287.0, f32)):
        v, dv = history_gather(lg, his, span, kind)
        v_r, d_r = oracle.history_gather(kind, his, span, lg)
        assert np.array_equal(v.cpu().numpy(), v_r)
        assert np.array_equal(dv.cpu().numpy(), d_r)
    gy = rng.standard_normal(v_r.shape).astype(f32)
    gl = history_gather_bwd(torch.from_numpy(gy).cuda(), dv)
    gl_r = oracle.history_gather_bwd(gy, d_r)
    np.testing.assert_allclose(gl.cpu().numpy(), gl_r, rtol=1e-5, atol=1e-5 * np.abs(gl_r).max())


def test_history_gather_nonuniform_span_and_reference_fixture(px, torch, oracle):
    """The reference's own interpolation fixtures (tests/interpolation/test_interpolation.py:13-85):
    ramp at t=21.12 and sin sampled at 0.01 queried at t=16.5."""
    ramp = np.arange(100, dtype=f32).reshape(1, 100, 1)
    for cls, kind in ((px.interpolation.LinearInterpolation, "linear"), (px.interpolation.CubicHermiteSpline, "cubic"),
                      (px.interpolation.BezierSpline, "bez")):
        it = cls(ramp, np.arange(100, dtype=f32))
        np.testing.assert_allclose(it.evaluate([21.12]).cpu().numpy().ravel(), [21.12], rtol=1e-4)
        np.testing.assert_allclose(it.derivative([21.12]).cpu().numpy().ravel(), [1.0], rtol=1e-4)
    ts = np.arange(0, 20, 0.01, dtype=f32)
    series = np.sin(ts).reshape(1, -1, 1).astype(f32)
    it = px.interpolation.CubicHermiteSpline(series, ts)
    np.testing.assert_allclose(it.evaluate([16.5]).cpu().numpy().ravel(), [np.sin(16.5)], rtol=1e-5)
    np.testing.assert_allclose(it.derivative([16.5]).cpu().numpy().ravel(), [np.cos(16.5)], rtol=1e-2)
    it = px.interpolation.BezierSpline(series, ts)  # test_interpolation.py:82-85
    np.testing.assert_allclose(it.evaluate([16.5]).cpu().numpy().ravel(), [np.sin(16.5)], rtol=5e-2)
    np.testing.assert_allclose(it.derivative([16.5]).cpu().numpy().ravel(), [np.cos(16.5)], rtol=1e-2)
    # non-uniform grid vs oracle
    rng = np.random.default_rng(9)
    span = np.cumsum(rng.uniform(0.2, 1.5, 40)).astype(f32)
    his = rng.standard_normal((5, 40, 7)).astype(f32)
    q = rng.uniform(span[0], span[-1], 33).astype(f32)
    from paddlexde_b200.xde.base_dde import history_gather

    for kind in ("linear", "cubic", "bez"):
        v, dv = history_gather(q, his, span, kind)
        v_r, d_r = oracle.history_gather(kind, his, span, q)
        assert np.array_equal(v.cpu().numpy(), v_r) and np.array_equal(dv.cpu().numpy(), d_r)
    with pytest.raises(ValueError):  # a Bezier segment needs four samples
        history_gather(q[:3], his[:, :3], span[:3], "bez")


def test_ddeint_one_damped_euler_step(px, torch, oracle):
    """D3STN usage (train_dde.py:418-433): lags -> y_lags, one Euler step with the damped fuse;
    gradient wrt the learnable lags through HistoryIndex.backward."""
    rng = np.random.default_rng(6)
    his = rng.uniform(-1, 1, (8, 307, 288, 3)).astype(f32)
    span = np.arange(288, dtype=f32)
    lags = torch.tensor((np.arange(12) + rng.uniform(0, 1, 12)).astype(f32), device="cuda", requires_grad=True)
    y0 = rng.uniform(-1, 1, (8, 307, 12, 3)).astype(f32)
    V = torch.tensor(rng.standard_normal((3, 3)).astype(f32) * 0.3, device="cuda")

    def func(y_lags, y):  # stand-in field; the real D3STN network is out of scope
        return torch.tanh(y @ V + y_lags.mean(dim=-2, keepdim=True))

    sol, y_lags = px.ddeint(func, torch.from_numpy(y0).cuda(), np.arange(2, dtype=f32), lags,
                            torch.from_numpy(his).cuda(), torch.from_numpy(span).cuda(), px.Euler,
                            fixed_solver_interp="")
    yl_ref, d_ref = oracle.history_gather("cubic", his, span, lags.detach().cpu().numpy())
    assert np.array_equal(y_lags.detach().cpu().numpy(), yl_ref)
    dy = func(torch.from_numpy(yl_ref).cuda(), torch.from_numpy(y0).cuda()).cpu().numpy()
    y1_ref = oracle.dde_fuse(dy, 1.0, y0)
    assert tuple(sol.shape) == (8, 307, 24, 3)  # concat(axis=-2) of y0 and y1
    assert np.array_equal(sol[..., 12:, :].detach().cpu().numpy(), y1_ref)
    gy = rng.standard_normal(yl_ref.shape).astype(f32)
    y_lags.backward(torch.from_numpy(gy).cuda())
    gl_ref = oracle.history_gather_bwd(gy, d_ref)
    np.testing.assert_allclose(lags.grad.cpu().numpy(), gl_ref, rtol=1e-5, atol=1e-5 * np.abs(gl_ref).max())
