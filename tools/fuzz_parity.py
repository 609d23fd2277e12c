"""Randomised parity sweep: the CUDA kernels (through the shim / C ABI) against the CPU oracle on random shapes,
weight scales, tolerances, controller options, time directions and batch sizes.  Bit-exact for states, step
counters and attempt logs; parameter gradients at rtol 1e-5 (+ an absolute floor under cancellation, see
case_adjoint).  Test infrastructure (it imports oracle/).

usage: python tools/fuzz_parity.py [seconds=120] [seed=0]      -> prints one line per failing case, then a summary
       python tools/fuzz_parity.py case SEED INDEX            -> re-runs one case (every case has its own generator)
"""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import paddlexde_b200 as px
from oracle import xde_oracle as xo
from paddlexde_b200.functional.odeint_adjoint import adjoint_backward

f32 = np.float32
ONE = len(sys.argv) > 1 and sys.argv[1] == "case"
budget = 0.0 if ONE else (float(sys.argv[1]) if len(sys.argv) > 1 else 120.0)
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(SEED)
xo.build()
PRES = ["id", "square", "cube"]
TILE = [(64, 256), (64, 128), (32, 256), (32, 128), (32, 64), (16, 64)]


def weights(d, h, scale):
    return [(scale * rng.standard_normal((d, h)) / np.sqrt(d)).astype(f32), (0.1 * rng.standard_normal(h)).astype(f32),
            (scale * rng.standard_normal((h, d)) / np.sqrt(h)).astype(f32), (0.1 * rng.standard_normal(d)).astype(f32)]


def tspan(n, span, reverse):
    t = np.sort(rng.uniform(0, span, n)).astype(f32)
    t = np.unique(t)
    if t.size < 2:
        t = np.array([0.0, span], f32)
    return t[::-1].copy() if reverse else t


def ctrl_opts():
    o = dict(rtol=float(10.0 ** rng.uniform(-7, -3)))
    o["atol"] = o["rtol"] * 1e-2
    if rng.random() < 0.3:
        o["first_step"] = float(10.0 ** rng.uniform(-4, -1))
    if rng.random() < 0.3:
        o["max_step"] = float(10.0 ** rng.uniform(-1.5, 0))
    if rng.random() < 0.3:
        o.update(safety=float(rng.uniform(0.7, 0.95)), ifactor=float(rng.uniform(3, 10)), dfactor=float(rng.uniform(0.1, 0.5)))
    return o


def case_dopri5_small():
    d = int(rng.integers(1, 9)); h = int(rng.integers(2, 70)); pre = PRES[rng.integers(3)]
    B = int(rng.integers(1, 400)); w = weights(d, h, rng.uniform(0.5, 3.0)); o = ctrl_opts()
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 9)), rng.uniform(0.2, 3), rng.random() < 0.3)
    desc = f"dopri5 small d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    xde = px.xde.BaseODE(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda(), t)
    s = px.Dopri5(xde=xde, y0=xde.y0, check_status=False, **o)
    sol = s.integrate(t).cpu().numpy()
    ref, st, _, rc = xo.dopri5_mlp(xo.MLP(*w, pre=pre), y0, t, **o)
    stt = s.read_stats()
    ok = np.array_equal(sol, ref, equal_nan=True) and stt.n_attempts == int(st.n_attempts.sum()) and stt.status == rc
    return desc, ok


def case_dopri5_tile():
    d, h = TILE[rng.integers(len(TILE))]; pre = PRES[rng.integers(3)]
    B = int(rng.integers(1, 200)); w = weights(d, h, rng.uniform(0.5, 3.0)); o = ctrl_opts()
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 7)), rng.uniform(0.2, 2), rng.random() < 0.3)
    desc = f"dopri5 tile d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    xde = px.xde.BaseODE(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda(), t)
    s = px.Dopri5(xde=xde, y0=xde.y0, check_status=False, log_attempts=256, **o)
    sol = s.integrate(t).cpu().numpy()
    om = xo.MLP(*w, pre=pre)
    ref, st, _, rc = xo.dopri5_mlp(om, y0, t, **o)
    stt = s.read_stats()
    ok = np.array_equal(sol, ref, equal_nan=True) and stt.n_attempts == int(st.n_attempts.sum()) and \
        stt.n_accepted == int(st.n_accepted.sum()) and stt.nfe == int(st.nfe.sum()) and stt.status == rc
    if ok:
        rec, cnt = s.attempt_log.read()
        b = int(rng.integers(B))
        _, _, lg, _ = xo.dopri5_mlp(om, y0, t, log_traj=b, **o)
        n = min(cnt[b], 256)
        ok = cnt[b] == len(lg) and rec[b, :n].tobytes() == lg[:n].tobytes()
    return desc, ok


def case_other_tableaux():
    name = ["Bosh3", "Fehlberg2", "AdaptiveHeun", "Dopri8"][rng.integers(4)]
    d = int(rng.integers(1, 9)); h = int(rng.integers(2, 50)); pre = PRES[rng.integers(3)]
    if name != "Dopri8" and rng.random() < 0.4:  # the tiled kernels (large states)
        d, h = TILE[rng.integers(len(TILE))]
    B = int(rng.integers(1, 200)); w = weights(d, h, rng.uniform(0.5, 2.0)); o = ctrl_opts()
    o["rtol"] = max(o["rtol"], 1e-6); o["atol"] = o["rtol"] * 1e-2
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 6)), rng.uniform(0.2, 1.5), rng.random() < 0.3)
    desc = f"{name} d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    xde = px.xde.BaseODE(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda(), t)
    s = getattr(px, name)(xde=xde, y0=xde.y0, check_status=False, **o)
    sol = s.integrate(t).cpu().numpy()
    key = {"Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "AdaptiveHeun": "adaptive_heun", "Dopri8": "dopri8"}[name]
    ref, st, _, rc = xo.adaptive_rk_mlp(key, xo.MLP(*w, pre=pre), y0, t, **o)
    stt = s.read_stats()
    return desc, np.array_equal(sol, ref, equal_nan=True) and stt.n_attempts == int(st.n_attempts.sum()) and stt.status == rc


def case_adjoint():
    d = int(rng.integers(1, 9)); h = int(rng.integers(2, 65)); pre = PRES[rng.integers(3)]
    B = int(rng.integers(1, 300)); w = weights(d, h, rng.uniform(0.5, 2.5))
    o = dict(rtol=float(10.0 ** rng.uniform(-7, -4))); o["atol"] = o["rtol"] * 1e-2
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 7)), rng.uniform(0.2, 2), rng.random() < 0.2)
    desc = f"adjoint d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    om = xo.MLP(*w, pre=pre)
    ref, _, _, rc = xo.dopri5_mlp(om, y0, t, **o)
    if rc != 0:
        return desc + " (forward status, skipped)", True
    gy = (rng.standard_normal(ref.shape) / ref[-1].size).astype(f32)
    want_gt = rng.random() < 0.4  # the t_requires_grad branch: grad_t_span, the g_t slot in the controller norm
    gt = torch.zeros(t.size, device="cuda") if want_gt else None
    gt_ref = np.zeros(t.size, f32) if want_gt else None
    desc += f" grad_t={want_gt}"
    g, a0, stats, _ = adjoint_backward(px.MLPField(*w, pre=pre), t, ref, gy, return_adj_y0=True, check_status=False,
                                       out_grad_t=gt, **o)
    g_ref, a_ref, st_ref, _, rc = xo.dopri5_mlp_adjoint(om, t, ref, gy, grad_t=gt_ref, **o)
    s = stats.read()
    scale = max(np.abs(g_ref).max(), 1e-30)
    checks = {"status": s.status == rc, "attempts": s.n_attempts == int(st_ref.n_attempts.sum()),
              "grad_t": (not want_gt) or rc != 0 or np.allclose(gt.cpu().numpy(), gt_ref, rtol=1e-5,
                                                                   atol=3e-5 * max(np.abs(gt_ref).max(), 1e-30)),
              "adj_state": np.array_equal(a0.cpu().numpy(), a_ref, equal_nan=True),
              # random-sign cotangents at every output time cancel across the batch, and the two sides sum the same
              # per-evaluation contributions in different fp32 orders (oracle: g_theta as fp32 Runge-Kutta state;
              # kernel: sum_i W_i k_i^theta folded per warp): the floor is absolute, 3e-5 of the largest gradient
              # (worst seen in 5 000 cases: 1.2e-5; flushing the kernel's fp32 partial sums 4x as often changes nothing)
              "grads": rc != 0 or np.allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=3e-5 * scale)}
    ok = all(checks.values())
    if not ok:
        gd = g.cpu().numpy()
        desc += f" failed={[k for k, v in checks.items() if not v]} status={s.status}/{rc} attempts={s.n_attempts}/" \
                f"{int(st_ref.n_attempts.sum())} max|dg|/max|g|={np.abs(gd - g_ref).max() / scale:.2e} " \
                f"n_adj_state_diff={int((a0.cpu().numpy() != a_ref).sum())}"
    return desc, ok


def case_adjoint_tile():
    """odeint_adjoint's backward for large states (csrc/xde_adj_tile.cu)"""
    d, h = TILE[rng.integers(len(TILE))]; pre = PRES[rng.integers(3)]
    B = int(rng.integers(1, 150)); w = weights(d, h, rng.uniform(0.5, 2.5))
    o = dict(rtol=float(10.0 ** rng.uniform(-6.5, -4))); o["atol"] = o["rtol"] * 1e-2
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 5)), rng.uniform(0.2, 1.5), rng.random() < 0.2)
    desc = f"adjoint tile d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    om = xo.MLP(*w, pre=pre)
    ref, _, _, rc = xo.dopri5_mlp(om, y0, t, **o)
    if rc != 0:
        return desc + " (forward status, skipped)", True
    gy = (rng.standard_normal(ref.shape) / ref[-1].size).astype(f32)
    g, a0, stats, _ = adjoint_backward(px.MLPField(*w, pre=pre), t, ref, gy, return_adj_y0=True, check_status=False, **o)
    g_ref, a_ref, st_ref, _, rc = xo.dopri5_mlp_adjoint(om, t, ref, gy, **o)
    s = stats.read()
    scale = max(np.abs(g_ref).max(), 1e-30)
    checks = {"status": s.status == rc, "attempts": s.n_attempts == int(st_ref.n_attempts.sum()),
              "accepted": s.n_accepted == int(st_ref.n_accepted.sum()),
              "adj_state": np.array_equal(a0.cpu().numpy(), a_ref, equal_nan=True),
              "grads": rc != 0 or np.allclose(g.cpu().numpy(), g_ref, rtol=1e-5, atol=3e-5 * scale)}
    ok = all(checks.values())
    if not ok:
        desc += f" failed={[k for k, v in checks.items() if not v]} max|dg|/max|g|={np.abs(g.cpu().numpy() - g_ref).max() / scale:.2e}"
    return desc, ok


def case_adjoint_batch():
    """controller='batch' (the reference's default configuration): the whole sequence and g_theta bit for bit"""
    d = [1, 2, 4][rng.integers(3)]; h = int(rng.integers(2, 65)); pre = PRES[rng.integers(3)]
    B = int(rng.integers(1, 3000)); w = weights(d, h, rng.uniform(0.5, 2.5))
    o = dict(rtol=float(10.0 ** rng.uniform(-7, -4))); o["atol"] = o["rtol"] * 1e-2
    norm = ["mixed", "seminorm"][rng.integers(2)]
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 6)), rng.uniform(0.2, 2), rng.random() < 0.2)
    desc = f"adjoint batch {norm} d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    om = xo.MLP(*w, pre=pre)
    ref, _, _, rc = xo.dopri5_mlp(om, y0, t, controller="batch", **o)
    if rc != 0:
        return desc + " (forward status, skipped)", True
    gy = (rng.standard_normal(ref.shape) / ref[-1].size).astype(f32)
    g, a0, stats, log = adjoint_backward(px.MLPField(*w, pre=pre), t, ref, gy, return_adj_y0=True, check_status=False,
                                         controller="batch", adj_norm=norm, log_attempts=2048, **o)
    g_ref, a_ref, st_ref, lg, rc = xo.dopri5_mlp_adjoint(om, t, ref, gy, controller="batch", adj_norm=norm, **o)
    s = stats.read()
    rec, cnt = log.read()
    n = min(int(cnt[0]), 2048)
    checks = {"status": s.status == rc, "log_len": int(cnt[0]) == len(lg),
              "sequence": rec[0, :n].tobytes() == lg[:n].tobytes(),
              "adj_state": rc != 0 or np.array_equal(a0.cpu().numpy(), a_ref, equal_nan=True),
              "g_theta": rc != 0 or np.array_equal(g.cpu().numpy(), g_ref, equal_nan=True)}
    ok = all(checks.values())
    if not ok:
        desc += f" failed={[k for k, v in checks.items() if not v]}"
    return desc, ok


def case_fixed_grid():
    """FixedSolver(step_size=): len(t) - 1 steps on the constructed grid + linear_interp at t[i]"""
    solver = ["Euler", "RK4", "Midpoint"][rng.integers(3)]
    d, h = (TILE[rng.integers(len(TILE))] if rng.random() < 0.3 else (int(rng.integers(1, 9)), int(rng.integers(2, 60))))
    pre = PRES[rng.integers(3)]; B = int(rng.integers(1, 200)); w = weights(d, h, rng.uniform(0.5, 2.0))
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    T = int(rng.integers(2, 8)); t = np.linspace(0, rng.uniform(0.3, 1.0), T).astype(f32)
    step = float((t[-1] - t[0]) / (T - 1) / rng.uniform(1.5, 6.0))
    niters = int(np.ceil(np.float32((t[-1] - t[0]) / np.float32(step)) + np.float32(1.0)))
    grid = np.arange(0, niters, dtype=f32) * np.float32(step) + t[0]
    grid[-1] = t[-1]
    desc = f"{solver} step_size={step:.4f} d={d} h={h} pre={pre} B={B} T={T}"
    g = np.ascontiguousarray(grid[:T])
    if not np.all(np.diff(g) > 0):
        return desc + " (degenerate grid, skipped)", True
    y = xo.fixed_mlp(solver.lower(), xo.MLP(*w, pre=pre), y0, g)
    ref = np.empty_like(y); ref[:, 0] = y[:, 0]
    for i in range(1, T):
        if t[i] == g[i - 1]: ref[:, i] = y[:, i - 1]
        elif t[i] == g[i]: ref[:, i] = y[:, i]
        else: ref[:, i] = y[:, i - 1] + np.float32(np.float32(t[i] - g[i - 1]) / np.float32(g[i] - g[i - 1])) * (y[:, i] - y[:, i - 1])
    sol = px.odeint(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda().reshape(B, 1, d), t, getattr(px, solver),
                    options={"math": "fp32", "step_size": step}).cpu().numpy()
    return desc, np.array_equal(sol, ref, equal_nan=True)


def case_fixed():
    solver = ["Euler", "RK4", "Midpoint"][rng.integers(3)]
    big = rng.random() < 0.4
    if big:
        d, h = TILE[rng.integers(len(TILE))] if rng.random() < 0.8 else (64, 64)
    else:
        d, h = int(rng.integers(1, 9)), int(rng.integers(2, 60))
    pre = PRES[rng.integers(3)]; B = int(rng.integers(1, 300)); w = weights(d, h, rng.uniform(0.5, 2.0))
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 12)), rng.uniform(0.2, 1.0), rng.random() < 0.3)
    desc = f"{solver} fp32 d={d} h={h} pre={pre} B={B} T={t.size}"
    sol = px.odeint(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda().reshape(B, 1, d), t, getattr(px, solver),
                    options={"math": "fp32"}).cpu().numpy()
    ref = xo.fixed_mlp(solver.lower(), xo.MLP(*w, pre=pre), y0, t)
    return desc, np.array_equal(sol, ref, equal_nan=True)


def case_sde():
    scheme = ["em", "milstein"][rng.integers(2)]
    d = int(rng.integers(1, 9)); h = int(rng.integers(2, 50)); B = int(rng.integers(1, 300))
    if rng.random() < 0.3:  # the FP32 tiles (Milstein included since round 2)
        d, h = [(32, 64), (32, 128), (16, 64), (64, 128), (64, 64)][rng.integers(5)]
    wf, wg = weights(d, h, 1.0), weights(d, h, 0.7)
    pf, pg = PRES[rng.integers(3)], PRES[rng.integers(3)]
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 10)), 1.0, False)
    dW = (0.2 * rng.standard_normal((t.size - 1, B, d))).astype(f32)
    desc = f"sde {scheme} d={d} h={h} pre={pf}/{pg} B={B} T={t.size}"
    sol = px.sdeint(px.MLPField(*wf, pre=pf), px.MLPField(*wg, pre=pg), torch.from_numpy(y0).cuda().reshape(B, 1, d), t,
                    px.Euler, options={"bm_increments": torch.from_numpy(dW).cuda(), "scheme": scheme, "math": "fp32"}).cpu().numpy()
    ref = xo.sde_mlp(scheme, xo.MLP(*wf, pre=pf), xo.MLP(*wg, pre=pg), y0, t, dW)
    return desc, np.array_equal(sol, ref, equal_nan=True)


def case_gather():
    from paddlexde_b200.xde.base_dde import history_gather, history_gather_bwd
    kind = ["linear", "cubic", "bez"][rng.integers(3)]
    Th = int(rng.integers(4, 40)); D = int(rng.integers(1, 9)); L = int(rng.integers(1, 30))
    lead = (int(rng.integers(1, 40)), int(rng.integers(1, 12)))
    span = np.cumsum(rng.uniform(0.2, 2.0, Th)).astype(f32) if rng.random() < 0.5 else np.arange(Th, dtype=f32)
    lags = rng.uniform(span[0] - 1.0, span[-1] + 1.0, L).astype(f32)
    if rng.random() < 0.5:  # queries exactly on grid points (right-open bucketize, interpolate_base.py:49-62)
        lags[: L // 2] = span[rng.integers(0, Th, L // 2)]
    his = rng.uniform(-2, 2, lead + (Th, D)).astype(f32)
    if rng.random() < 0.2:
        his[..., 1:] = np.round(his[..., 1:] * 5)  # integer-valued channels like the dataset
    desc = f"gather {kind} lead={lead} Th={Th} D={D} L={L}"
    v, dv = history_gather(torch.from_numpy(lags).cuda(), torch.from_numpy(his).cuda(), torch.from_numpy(span).cuda(), kind)
    v_ref, d_ref = xo.history_gather(kind, his, span, lags)
    ok = np.array_equal(v.cpu().numpy(), v_ref, equal_nan=True) and np.array_equal(dv.cpu().numpy(), d_ref, equal_nan=True)
    if ok:
        gy = rng.standard_normal(v_ref.shape).astype(f32)
        gl = history_gather_bwd(torch.from_numpy(gy).cuda(), dv).cpu().numpy()
        gl_ref = xo.history_gather_bwd(gy, d_ref)
        ok = np.allclose(gl, gl_ref, rtol=1e-5, atol=1e-5 * max(np.abs(gl_ref).max(), 1e-30))
        desc += " (bwd)"
    return desc, ok


def case_batch_controller():
    d = int(rng.integers(1, 9)); h = int(rng.integers(2, 60)); pre = PRES[rng.integers(3)]
    B = int(rng.integers(1, 3000)); w = weights(d, h, rng.uniform(0.5, 2.5)); o = ctrl_opts()
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 7)), rng.uniform(0.2, 2), rng.random() < 0.3)
    desc = f"dopri5 controller=batch d={d} h={h} pre={pre} B={B} T={t.size} {o}"
    xde = px.xde.BaseODE(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda(), t)
    s = px.Dopri5(xde=xde, y0=xde.y0, check_status=False, controller="batch", **o)
    sol = s.integrate(t).cpu().numpy()
    ref, st, lg, rc = xo.dopri5_mlp(xo.MLP(*w, pre=pre), y0, t, controller="batch", **o)
    stt = s.read_stats()
    return desc, np.array_equal(sol, ref, equal_nan=True) and stt.status == rc


def case_grid_points():
    name = ["Dopri5", "Bosh3", "Fehlberg2", "AdaptiveHeun", "Dopri8"][rng.integers(5)]
    d = int(rng.integers(1, 9)); h = int(rng.integers(2, 40)); pre = PRES[rng.integers(3)]
    if name != "Dopri8" and rng.random() < 0.4:  # the tiled kernels (large states)
        d, h = TILE[rng.integers(len(TILE))]
    B = int(rng.integers(1, 150)); w = weights(d, h, rng.uniform(0.5, 2.0))
    o = dict(rtol=float(10.0 ** rng.uniform(-6, -3))); o["atol"] = o["rtol"] * 1e-2
    rev = rng.random() < 0.3
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 6)), 1.5, rev)
    step_t = [float(x) for x in rng.uniform(-0.2, 1.7, int(rng.integers(0, 5)))]
    jump_t = [float(x) for x in rng.uniform(-0.2, 1.7, int(rng.integers(0, 4)))]
    if not step_t and not jump_t:
        step_t = [0.5]
    desc = f"{name} step_t={np.round(step_t, 3)} jump_t={np.round(jump_t, 3)} d={d} h={h} pre={pre} B={B} T={t.size} rev={rev} {o}"
    xde = px.xde.BaseODE(px.MLPField(*w, pre=pre), torch.from_numpy(y0).cuda(), t)
    s = getattr(px, name)(xde=xde, y0=xde.y0, check_status=False, step_t=step_t or None, jump_t=jump_t or None, **o)
    sol = s.integrate(t).cpu().numpy()
    key = {"Dopri5": "dopri5", "Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "AdaptiveHeun": "adaptive_heun", "Dopri8": "dopri8"}[name]
    ref, st, _, rc = xo.adaptive_rk_mlp(key, xo.MLP(*w, pre=pre), y0, t, step_t=step_t or None, jump_t=jump_t or None, **o)
    stt = s.read_stats()
    return desc, np.array_equal(sol, ref, equal_nan=True) and stt.n_attempts == int(st.n_attempts.sum()) and \
        stt.nfe == int(st.nfe.sum()) and stt.status == rc


def case_sde_adjoint():
    from paddlexde_b200.functional.sdeint_adjoint import sde_adjoint_backward
    d = [1, 2, 3, 4, 8][rng.integers(5)]; h = int(rng.integers(2, 90)); B = int(rng.integers(1, 300))
    wf, wg = weights(d, h, 1.0), weights(d, h, 0.7)
    pf, pg = PRES[rng.integers(3)], PRES[rng.integers(3)]
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32); t = tspan(int(rng.integers(2, 10)), 1.0, False)
    dW = (0.2 * rng.standard_normal((t.size - 1, B, d))).astype(f32)
    desc = f"sde adjoint d={d} h={h} pre={pf}/{pg} B={B} T={t.size}"
    f, g = px.MLPField(*wf, pre=pf), px.MLPField(*wg, pre=pg)
    of, og = xo.MLP(*wf, pre=pf), xo.MLP(*wg, pre=pg)
    table = torch.from_numpy(dW).cuda()
    sol = px.sdeint(f, g, torch.from_numpy(y0).cuda().reshape(B, 1, d), t, px.Euler, options={"bm_increments": table})
    ref = xo.sde_mlp("em", of, og, y0, t, dW)
    if not np.array_equal(sol.cpu().numpy(), ref, equal_nan=True):
        return desc + " (forward)", False
    gy = (np.abs(rng.standard_normal(ref.shape)) / ref.size).astype(f32)
    gf_r, gg_r, a0_r = xo.sde_mlp_adjoint(of, og, t, ref, gy, dW)
    gf, gg, a0 = sde_adjoint_backward(f, g, t, sol, torch.from_numpy(gy).cuda(), bm_increments=table, return_adj_y0=True)
    ok = np.array_equal(a0.cpu().numpy(), a0_r, equal_nan=True)
    for got, want in ((gf, gf_r), (gg, gg_r)):
        ok = ok and np.allclose(got.cpu().numpy(), want, rtol=1e-5, atol=3e-5 * max(np.abs(want).max(), 1e-30))
    return desc, ok


def case_tensor():
    """tcgen05 path vs the FP32 kernels (themselves bit-exact vs the oracle): ragged batches, many tiles per CTA,
    strides, supplied and generated increments; 1e-5 of the largest state."""
    shapes = TILE + [(64, 64)]
    kind = ["Euler", "RK4", "Midpoint", "sde", "sde_gen"][rng.integers(5)]
    d, h = shapes[rng.integers(len(shapes))]
    if kind.startswith("sde") and (d == 64 and h > 64 or h > 128):
        d, h = 32, 64  # TMEM limits of the SDE variant (two networks)
    B = int(rng.integers(1, 70000)) if rng.random() < 0.3 else int(rng.integers(1, 700))
    T = int(rng.integers(2, 8)); stride = int(rng.integers(1, 4))
    t = np.linspace(0, rng.uniform(0.2, 1.0), T).astype(f32)
    y0 = torch.from_numpy(rng.uniform(-1, 1, (B, 1, d)).astype(f32)).cuda()
    desc = f"tensor {kind} d={d} h={h} B={B} T={T} stride={stride}"
    if not kind.startswith("sde"):
        pre = PRES[rng.integers(3)]
        field = px.MLPField(*weights(d, h, rng.uniform(0.5, 1.5)), pre=pre)
        desc += f" pre={pre}"
        run = lambda math: px.odeint(field, y0, t, getattr(px, kind), options={"math": math, "out_stride": stride})
    else:
        f = px.MLPField(*weights(d, h, 1.0), pre=PRES[rng.integers(3)])
        g = px.MLPField(*weights(d, h, 0.7), pre=PRES[rng.integers(3)])
        if kind == "sde":
            dW = torch.randn((T - 1, B, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))) * 0.2
            opt = {"bm_increments": dW}
        else:
            opt = {"bm_seed": int(rng.integers(1 << 30))}
        run = lambda math: px.sdeint(f, g, y0, t, px.Euler, options={**opt, "math": math, "out_stride": stride})
    a, b = run("tensor"), run("fp32")
    scale = float(b.abs().max())
    err = float((a - b).abs().max())
    ok = a.shape == b.shape and err <= 1e-5 * scale
    if not ok:
        desc += f" max|d|={err:.3e} scale={scale:.3e}"
    return desc, ok


CASES = [case_dopri5_small, case_dopri5_tile, case_other_tableaux, case_adjoint, case_fixed, case_sde, case_gather,
         case_batch_controller, case_grid_points, case_sde_adjoint, case_tensor, case_adjoint_tile, case_adjoint_batch,
         case_fixed_grid]
counts = {c.__name__: [0, 0] for c in CASES}
t_end = time.time() + budget
i = int(sys.argv[3]) if ONE else 0
while ONE or time.time() < t_end:
    c = CASES[i % len(CASES)]
    rng = np.random.default_rng([SEED, i])
    i += 1
    try:
        desc, ok = c()
    except Exception as e:  # an exception is a finding too
        desc, ok = f"{c.__name__}: {type(e).__name__}: {e}", False
        traceback.print_exc()
    counts[c.__name__][0] += 1
    if not ok:
        counts[c.__name__][1] += 1
        print(f"MISMATCH (case {SEED} {i - 1}):", desc, flush=True)
    if ONE:
        print("ok" if ok else "FAILED", desc)
        break
print("cases run / failed:", {k: tuple(v) for k, v in counts.items()})
