"""C oracle (fused-MLP specialisation) vs the literal NumPy restatement and vs fp64 truths.

No reference test pins any of this (SURVEY.md 8(c) "not pinned"): step sequences, adjoint gradients,
SDE results and B>1 behaviour are defined by the restated oracle -- "parity unpinned"."""
import numpy as np
import pytest

from oracle import oracle_np as onp
from tests.problems import cfg2_tspan, cfg2_y0, fanin_weights, spiral_weights

f32 = np.float32


@pytest.fixture(scope="module")
def spiral(oracle):
    return oracle.MLP(*spiral_weights(), pre="cube")


def test_field_close_to_fp64(oracle, spiral):
    y = cfg2_y0(64)
    ref = onp.MLPFieldNP(spiral.w1, spiral.b1, spiral.w2, spiral.b2, "cube", np.float64)(0.0, y)
    np.testing.assert_allclose(spiral(0.0, y), ref, rtol=2e-5, atol=2e-6)
    c = np.random.default_rng(2).standard_normal(y.shape).astype(f32)
    f, dy, gs = spiral.vjp(0.0, y, c)
    f64 = onp.MLPFieldNP(spiral.w1, spiral.b1, spiral.w2, spiral.b2, "cube", np.float64)
    f_r, dy_r, gs_r = f64.vjp(0.0, y, c)
    np.testing.assert_allclose(dy, dy_r, rtol=1e-4, atol=1e-5)
    for g, gr in zip(gs, gs_r):
        np.testing.assert_allclose(g, gr, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("controller", ["trajectory", "batch"])
def test_dopri5_c_equals_literal(oracle, spiral, controller):
    y0, t = cfg2_y0(12), np.linspace(0, 25, 1000).astype(f32)[:40:4]
    out, st, log, rc = oracle.dopri5_mlp(spiral, y0, t, controller=controller, log_traj=5)
    assert rc == 0
    if controller == "batch":
        sol = onp.odeint(spiral, y0, t, onp.Dopri5)
        assert np.array_equal(sol, out)
    else:
        sol = onp.odeint(spiral, y0[5:6], t, onp.Dopri5)
        assert np.array_equal(sol[:, 0], out[:, 5])
    lg = onp.odeint.last_log
    assert np.array_equal(np.array(lg.dt, f32), log.dt)
    assert np.array_equal(np.array(lg.ratio, f32), log.ratio)
    assert lg.accepted == list(log.accepted.astype(bool))
    assert lg.nfe == (st.nfe[5] if controller == "trajectory" else st.nfe[0])


def test_dopri5_rejections_and_tolerance(oracle):
    # a stiffer field (fan-in weights, identity pre-activation) so the controller both accepts and rejects
    m = oracle.MLP(*[3.0 * a for a in fanin_weights(2, 50, seed=5)], pre="id")
    y0 = np.random.default_rng(1).uniform(-1, 1, (8, 2)).astype(f32)
    t = np.linspace(0, 4, 9).astype(f32)
    out, st, log, rc = oracle.dopri5_mlp(m, y0, t, log_traj=0, rtol=1e-6, atol=1e-8)
    assert rc == 0 and (st.n_attempts > st.n_accepted).any()
    sol = onp.odeint(m, y0[0:1], t, onp.Dopri5, rtol=1e-6, atol=1e-8)
    assert np.array_equal(sol[:, 0], out[:, 0])
    # accuracy vs scipy fp64
    from scipy.integrate import solve_ivp

    f64 = onp.MLPFieldNP(m.w1, m.b1, m.w2, m.b2, "id", np.float64)
    ref = solve_ivp(lambda tt, y: f64(tt, y), (0, 4), y0[0].astype(np.float64), t_eval=t.astype(np.float64),
                    rtol=1e-11, atol=1e-13, method="DOP853").y.T
    np.testing.assert_allclose(out[:, 0], ref, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("method,cls,rtol", [("bosh3", onp.Bosh3, 1e-6), ("fehlberg2", onp.Fehlberg2, 1e-4),
                                             ("adaptive_heun", onp.AdaptiveHeun, 1e-4),
                                             ("dopri8", onp.Dopri8, 1e-7), ("dopri5", onp.Dopri5, 1e-7)])
def test_other_tableaux_c_equals_literal(oracle, spiral, method, cls, rtol):
    """The C driver with each embedded tableau == the literal NumPy restatement, bit for bit (results
    and the whole accept/reject log), per-trajectory controller."""
    y0 = np.array([[2.0, 0.0], [1.5, -0.5], [-0.3, 1.1]], np.float32)
    t = np.linspace(0, 1.5, 7).astype(np.float32)
    out, st, log, rc = oracle.adaptive_rk_mlp(method, spiral, y0, t, log_traj=1, rtol=rtol, atol=rtol * 1e-2)
    assert rc == 0
    sol = onp.odeint(spiral, y0[1:2], t, cls, rtol=rtol, atol=rtol * 1e-2)
    lg = onp.odeint.last_log
    assert np.array_equal(out[:, 1:2], sol)
    assert len(lg.dt) == len(log) and np.array_equal(np.array(lg.dt, np.float32), log.dt)
    assert np.array_equal(np.array(lg.accepted), log.accepted.astype(bool))
    assert st.n_attempts[1] == len(log) and st.nfe[1] == lg.nfe


@pytest.mark.parametrize("p", [2, 3, 5, 8])
def test_rootp(oracle, p):
    r = np.exp(np.random.default_rng(p).uniform(-30, 30, 2000)).astype(np.float32)
    ref = r.astype(np.float64) ** (1.0 / p)
    got = oracle.rootpf(r, p)
    assert np.max(np.abs(got - ref) / ref) < 3e-7
    assert all(onp.rootp(v, p) == c for v, c in zip(r[:200], got[:200]))


@pytest.mark.parametrize("method,cls", [("euler", onp.Euler), ("rk4", onp.RK4), ("midpoint", onp.Midpoint)])
def test_fixed_c_equals_literal(oracle, spiral, method, cls):
    y0, t = cfg2_y0(6), np.linspace(0, 25, 1000).astype(f32)[:32]
    out = oracle.fixed_mlp(method, spiral, y0, t)                 # [B,T,D]
    sol = onp.odeint(spiral, y0[:, None, :], t, cls)              # [B,T,D]
    assert np.array_equal(out, sol)


def test_adjoint_c_equals_literal_and_fp64(oracle, spiral):
    import torch

    B = 6
    y0, t = cfg2_y0(B), cfg2_tspan(6)
    out, _, _, rc = oracle.dopri5_mlp(spiral, y0, t)
    gy = np.zeros_like(out)
    gy[-1] = np.sign(out[-1]) / out[-1].size
    gy[2] = 0.01                                                   # a mid-trajectory loss term too
    g, a0, st, log, rc = oracle.dopri5_mlp_adjoint(spiral, t, out, gy, log_traj=2)
    assert rc == 0
    gs = np.zeros(spiral.n_params, np.float64)
    for b in range(B):
        ps, a, logs = onp.odeint_adjoint_backward(spiral, t, out[:, b:b + 1], gy[:, b:b + 1], seminorm=True)
        gs += np.concatenate([p.ravel() for p in ps]).astype(np.float64)
        assert np.array_equal(a[0], a0[b])
        if b == 2:
            dts = np.concatenate([np.array(l.dt, f32) for l in logs])
            assert np.array_equal(dts, log.dt)
    assert np.array_equal(gs.astype(f32), g)
    # batch controller + default mixed norm (the literal reference configuration)
    gb, a0b, stb, logb, rc = oracle.dopri5_mlp_adjoint(spiral, t, out, gy, controller="batch", adj_norm="mixed")
    ps, a, logs = onp.odeint_adjoint_backward(spiral, t, out, gy, seminorm=False)
    np.testing.assert_allclose(np.concatenate([p.ravel() for p in ps]), gb, rtol=1e-5, atol=1e-8)
    # fp64 truth by autograd through a fine RK4
    W = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (spiral.w1, spiral.b1, spiral.w2, spiral.b2)]
    y = torch.tensor(y0, dtype=torch.float64)
    ts = t.astype(np.float64)
    loss = 0.0
    for i in range(1, len(ts)):
        n = 200
        h = (ts[i] - ts[i - 1]) / n
        fn = lambda y: torch.tanh((y ** 3) @ W[0] + W[1]) @ W[2] + W[3]
        for _ in range(n):
            k1 = fn(y); k2 = fn(y + h / 2 * k1); k3 = fn(y + h / 2 * k2); k4 = fn(y + h * k3)
            y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        loss = loss + (y * torch.tensor(gy[i], dtype=torch.float64)).sum()
    loss.backward()
    gt = np.concatenate([w.grad.numpy().ravel() for w in W])
    scale = np.max(np.abs(gt))
    assert np.max(np.abs(g - gt)) / scale < 5e-6
    assert np.max(np.abs(gb - gt)) / scale < 5e-6


def test_sde_em_and_milstein(oracle):
    D, H, B, T = 4, 16, 10, 9
    drift = oracle.MLP(*fanin_weights(D, H, seed=2), pre="cube")
    diff = oracle.MLP(*fanin_weights(D, H, seed=3), pre="square")
    rng = np.random.default_rng(2)
    y0 = rng.uniform(-1, 1, (B, D)).astype(f32)
    t = np.linspace(0, 1, T).astype(f32)
    dW = (np.sqrt(1 / (T - 1)) * rng.standard_normal((T - 1, B, D))).astype(f32)
    out = oracle.sde_mlp("em", drift, diff, y0, t, dW)
    # literal: Euler step with move/fuse of the intended BaseSDE (xde/base_sde.py:44-61)
    y = y0.copy()
    for n in range(T - 1):
        dt = f32(t[n + 1] - t[n])
        y = ((y + drift(t[n], y) * dt).astype(f32) + (diff(t[n], y) * dW[n]).astype(f32)).astype(f32)
        assert np.array_equal(out[:, n + 1], y)
    # Milstein (extension, no reference counterpart): check the analytic diagonal Jacobian by differences
    outm = oracle.sde_mlp("milstein", drift, diff, y0, t, dW)
    g64 = onp.MLPFieldNP(diff.w1, diff.b1, diff.w2, diff.b2, "square", np.float64)
    y = y0[:, :].astype(np.float64)
    eps = 1e-6
    gp = np.stack([(g64(0, y + eps * np.eye(D)[d])[:, d] - g64(0, y - eps * np.eye(D)[d])[:, d]) / (2 * eps)
                   for d in range(D)], axis=1)
    dt = float(t[1] - t[0])
    g = g64(0, y)
    f64 = onp.MLPFieldNP(drift.w1, drift.b1, drift.w2, drift.b2, "cube", np.float64)
    ref = y + f64(0, y) * dt + g * dW[0] + 0.5 * g * gp * (dW[0].astype(np.float64) ** 2 - dt)
    np.testing.assert_allclose(outm[:, 1], ref, rtol=2e-5, atol=2e-6)


def test_philox_known_answers_and_increment_layout():
    """Random123 known-answer vectors for Philox4x32-10 pin the generator restatement (oracle/philox_np.py)
    that the device-side Brownian increments are checked against."""
    from oracle.philox_np import brownian_increments, philox4x32_10

    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in philox4x32_10(*ctr, *key)) == want
    t = np.linspace(0, 1, 5).astype(np.float32)
    a = brownian_increments(7, t, 64, 6)
    assert a.shape == (4, 64, 6) and np.isfinite(a).all()
    # counter-based: a shard that passes its global offset sees the increments of the whole batch
    assert np.array_equal(brownian_increments(7, t, 24, 6, offset=40), a[:, 40:])
    assert not np.array_equal(brownian_increments(8, t, 64, 6), a)
    big = brownian_increments(3, np.array([0, 0.25], np.float32), 40000, 8)
    assert abs(big.mean()) < 3e-3 and abs(big.var() - 0.25) < 3e-3


def test_sde_adjoint_oracle_is_the_gradient_of_the_em_recursion(oracle):
    """orc_sde_mlp_adjoint (parity unpinned: the reference's sdeint_adjoint backward is a placeholder) against
    fp64 autograd through the same Euler-Maruyama recursion."""
    import torch
    from tests.problems import fanin_weights

    d, h, B, T = 4, 33, 29, 9
    wf, wg = fanin_weights(d, h, seed=2), fanin_weights(d, h, seed=3)
    of, og = oracle.MLP(*wf, pre="cube"), oracle.MLP(*wg, pre="square")
    rng = np.random.default_rng(0)
    y0 = rng.uniform(-1, 1, (B, d)).astype(np.float32)
    t = np.linspace(0, 1, T).astype(np.float32)
    dW = (np.sqrt(1 / (T - 1)) * rng.standard_normal((T - 1, B, d))).astype(np.float32)
    sol = oracle.sde_mlp("em", of, og, y0, t, dW)
    gy = (rng.standard_normal(sol.shape) / sol.size).astype(np.float32)
    gf, gg, a0 = oracle.sde_mlp_adjoint(of, og, t, sol, gy, dW)
    P = [torch.tensor(np.asarray(a, np.float64), requires_grad=True) for a in (*wf, *wg)]
    F = lambda y, w1, b1, w2, b2, p: torch.tanh((y ** p) @ w1 + b1) @ w2 + b2
    y = torch.tensor(y0.astype(np.float64), requires_grad=True)
    ys = [y]
    for n in range(T - 1):
        y = y + F(y, *P[:4], 3) * (float(t[n + 1]) - float(t[n])) + F(y, *P[4:], 2) * torch.tensor(dW[n].astype(np.float64))
        ys.append(y)
    (torch.stack(ys, 1) * torch.tensor(gy.astype(np.float64))).sum().backward()
    ref_f = np.concatenate([p.grad.numpy().ravel() for p in P[:4]])
    ref_g = np.concatenate([p.grad.numpy().ravel() for p in P[4:]])
    for got, want in ((gf, ref_f), (gg, ref_g), (a0, ys[0].grad.numpy())):
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_oracle_grad_t_span_is_the_true_time_gradient(oracle):
    """t_requires_grad branch of the adjoint (functional/odeint_adjoint.py:129-141,161-162) in the C oracle, against
    fp64 autograd through a fine RK4 integration whose step sizes are differentiable functions of t_span."""
    import torch

    rng = np.random.default_rng(8)
    d, h, B = 3, 12, 5
    w = [rng.standard_normal((d, h)) / np.sqrt(d), 0.1 * rng.standard_normal(h),
         rng.standard_normal((h, d)) / np.sqrt(h), 0.1 * rng.standard_normal(d)]
    om = oracle.MLP(*[a.astype(np.float32) for a in w], pre="id")
    y0 = rng.uniform(-1, 1, (B, d)).astype(np.float32)
    t = np.array([0.0, 0.3, 0.7, 1.2], np.float32)
    ref, _, _, rc = oracle.dopri5_mlp(om, y0, t)
    assert rc == 0
    gy = rng.standard_normal(ref.shape).astype(np.float32)
    gt = np.zeros(t.size, np.float32)
    g, a0, _, _, rc = oracle.dopri5_mlp_adjoint(om, t, ref, gy, grad_t=gt)
    assert rc == 0
    g_plain, a_plain, _, _, _ = oracle.dopri5_mlp_adjoint(om, t, ref, gy)
    np.testing.assert_allclose(a0, a_plain, rtol=1e-5, atol=1e-6)  # the g_t slot only perturbs the step sequence

    W = [torch.tensor(np.asarray(a, np.float32), dtype=torch.float64) for a in w]
    f = lambda y: torch.tanh(y @ W[0] + W[1]) @ W[2] + W[3]
    tt = torch.tensor(t, dtype=torch.float64, requires_grad=True)
    y = torch.tensor(y0, dtype=torch.float64)
    sols = [y]
    for i in range(1, t.size):
        n = 64
        hh = (tt[i] - tt[i - 1]) / n
        for _ in range(n):
            k1 = f(y); k2 = f(y + 0.5 * hh * k1); k3 = f(y + 0.5 * hh * k2); k4 = f(y + hh * k3)
            y = y + hh / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        sols.append(y)
    L = (torch.stack(sols) * torch.tensor(gy, dtype=torch.float64)).sum()
    L.backward()
    want = tt.grad.numpy()
    np.testing.assert_allclose(gt, want, rtol=2e-4, atol=2e-5 * np.abs(want).max())


def test_oracle_exact_batch_sum_is_order_independent_and_correctly_rounded(oracle):
    """fx_add_float / fx_to_float (the order-independent batch sum of the arithmetic specification): any permutation
    gives the same bits, and the result is the correctly rounded exact sum (addends above 2^-35 are not truncated)."""
    import ctypes as C
    from fractions import Fraction

    lib = oracle.lib()
    lib.orc_fx_sum.restype = C.c_float
    lib.orc_fx_sum.argtypes = [C.c_void_p, C.c_int64]

    def fx(x):
        x = np.ascontiguousarray(x, np.float32)
        return np.float32(lib.orc_fx_sum(x.ctypes.data_as(C.c_void_p), x.size))

    def rn32(fr):  # exact rational -> fp32, round to nearest even
        c = np.float32(float(fr))
        cands = sorted({np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))}, key=float)
        best = min(cands, key=lambda v: (abs(Fraction(float(v)) - fr), int(np.float32(v).view(np.uint32)) & 1))
        return np.float32(best)

    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(1, 4000))
        x = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 6, n)).astype(np.float32)
        if trial % 3 == 0:  # heavy cancellation
            x = np.concatenate([x, -x[: n // 2], (1e-3 * rng.standard_normal(5)).astype(np.float32)])
        want = rn32(sum((Fraction(float(v)) for v in x), Fraction(0)))
        got = fx(x)
        assert got == want, (trial, got, want)
        for _ in range(3):
            assert fx(rng.permutation(x)) == got
    # ties round to even; tiny addends are truncated toward zero to the 2^-59 grid; unrepresentable addends -> NaN
    assert fx([np.float32(2 ** 24), np.float32(1.0)]) == np.float32(2 ** 24)
    assert fx([np.float32(2 ** 24), np.float32(3.0)]) == np.float32(2 ** 24 + 4)
    assert fx([np.float32(2.0 ** -60), np.float32(-2.0 ** -60)]) == 0.0 and fx([np.float32(2.0 ** -61)] * 8) == 0.0
    assert fx([np.float32(3 * 2.0 ** -60)]) == np.float32(2.0 ** -59)
    assert np.isnan(fx([np.float32(2.0 ** 40)])) and np.isnan(fx([np.float32(np.inf), 1.0])) and fx([np.float32(2.0 ** 39)]) == 2.0 ** 39


def test_batch_adjoint_c_equals_an_independent_restatement_of_the_exact_sum(oracle):
    """The order-independent batch sum of the arithmetic specification (round 2), restated independently in exact
    rational arithmetic: 32-trajectory fp32 fma chains (fma = one rounding of the exact a*b + c), addends truncated toward
    zero to the 2^-59 grid, added exactly, one rounding to fp32.  Driving the literal NumPy solver (oracle_np) with this
    VJP must reproduce the C oracle's batch-controller adjoint -- the reference's default mixed-norm configuration --
    bit for bit: gradients, dL/dy0 and every dt of the step sequence."""
    import ctypes as C
    from fractions import Fraction

    def rn32(fr):  # exact rational -> fp32, round to nearest even
        if fr == 0:
            return f32(0.0)
        c = f32(float(fr))
        cands = sorted({np.nextafter(c, f32(-np.inf)), c, np.nextafter(c, f32(np.inf))}, key=float)
        return f32(min(cands, key=lambda v: (abs(Fraction(float(v)) - fr), int(f32(v).view(np.uint32)) & 1)))

    def fma32(a, b, c):
        return rn32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))

    GRID = 2 ** 59

    def exact_total(addends):
        tot = 0
        for x in addends:
            fr = Fraction(float(x)) * GRID
            tot += int(fr) if fr >= 0 else -int(-fr)  # truncation toward zero
        return rn32(Fraction(tot, GRID))

    rng = np.random.default_rng(11)
    d, h, B = 2, 6, 40  # two chain blocks: 32 + 8 trajectories
    w = [(rng.standard_normal((d, h)) / np.sqrt(d)).astype(f32), (0.1 * rng.standard_normal(h)).astype(f32),
         (rng.standard_normal((h, d)) / np.sqrt(h)).astype(f32), (0.1 * rng.standard_normal(d)).astype(f32)]
    mlp = oracle.MLP(*w, pre="id")
    lib = oracle.lib()

    def vjp_exact(t, y, c):
        y, c = np.ascontiguousarray(y, f32).reshape(-1, d), np.ascontiguousarray(c, f32).reshape(-1, d)
        n = y.shape[0]
        f, dy = np.empty_like(y), np.empty_like(y)
        per = []  # per trajectory: (u, dz, h, cot) from the C field evaluation (that part is covered by other tests)
        m = mlp.c()
        for b in range(n):
            g1 = np.zeros(mlp.n_params, f32)
            gw1, gb1, gw2, gb2 = mlp.split(g1)
            lib.orc_mlp_vjp(C.byref(m), y[b].ctypes.data_as(C.c_void_p), c[b].ctypes.data_as(C.c_void_p),
                            f[b].ctypes.data_as(C.c_void_p), dy[b].ctypes.data_as(C.c_void_p),
                            gw1.ctypes.data_as(C.c_void_p), gb1.ctypes.data_as(C.c_void_p),
                            gw2.ctypes.data_as(C.c_void_p), gb2.ctypes.data_as(C.c_void_p))
            hb = np.empty(h, f32)
            ftmp = np.empty(d, f32)
            lib.orc_mlp_eval(C.byref(m), y[b].ctypes.data_as(C.c_void_p), ftmp.ctypes.data_as(C.c_void_p),
                             hb.ctypes.data_as(C.c_void_p))
            per.append((y[b].copy(), gb1.copy(), hb, c[b].copy()))  # pre = id: u = y;  gb1 of one trajectory = dz
        if n == 1:
            g1 = np.zeros(mlp.n_params, f32)
            u, dz, hb, cot = per[0]
            return f, dy, [np.outer(u, dz).astype(f32), dz, np.outer(hb, cot).astype(f32), cot]
        gW1, gb1t, gW2, gb2t = np.zeros((d, h), f32), np.zeros(h, f32), np.zeros((h, d), f32), np.zeros(d, f32)
        blocks = [per[i:i + 32] for i in range(0, n, 32)]
        for k in range(d):
            for j in range(h):
                chains = []
                for blk in blocks:
                    X = f32(0.0)
                    for (u, dz, hb, cot) in blk:
                        X = fma32(u[k], dz[j], X)
                    chains.append(X)
                gW1[k, j] = exact_total(chains)
        for j in range(h):
            chains = []
            for blk in blocks:
                X = f32(0.0)
                for (u, dz, hb, cot) in blk:
                    X = fma32(f32(1.0), dz[j], X)
                chains.append(X)
            gb1t[j] = exact_total(chains)
            for dd in range(d):
                chains = []
                for blk in blocks:
                    X = f32(0.0)
                    for (u, dz, hb, cot) in blk:
                        X = fma32(cot[dd], hb[j], X)
                    chains.append(X)
                gW2[j, dd] = exact_total(chains)
        for dd in range(d):
            gb2t[dd] = exact_total([cot[dd] for (_, _, _, cot) in per])
        return f, dy, [gW1, gb1t, gW2, gb2t]

    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    t = np.array([0.0, 0.4, 1.0], f32)
    kw = dict(rtol=1e-6, atol=1e-8)
    out, _, _, rc = oracle.dopri5_mlp(mlp, y0, t, controller="batch", **kw)
    assert rc == 0
    gy = (rng.standard_normal(out.shape) / out[0].size).astype(f32)
    g, a0, st, log, rc = oracle.dopri5_mlp_adjoint(mlp, t, out, gy, controller="batch", adj_norm="mixed", **kw)
    assert rc == 0
    ps, a, logs = onp.odeint_adjoint_backward(mlp, t, out, gy, seminorm=False, vjp=vjp_exact, **kw)
    assert np.array_equal(np.concatenate([p.ravel() for p in ps]), g)
    assert np.array_equal(a, a0)
    assert np.array_equal(np.concatenate([np.array(l.dt, f32) for l in logs]), log.dt)
