"""Golden vectors of `odeint_adjoint(...)` + `backward` produced by the REFERENCE'S OWN CODE: the unmodified
`paddlexde/functional/odeint_adjoint.py` (option defaulting, `handle_adjoint_norm_`, `OdeintAdjointMethod.forward` and
`.backward` with `augmented_dynamics`, the segment loop and the `t_requires_grad` branch) driving the reference's
unmodified Dopri5 (`solver/adaptive_solver/dopri5.py` on `base_adaptive_solver_rk.py`) on the NumPy `paddle` stand-in of
oracle/ref_shim.

    python tools/make_reference_adjoint_golden.py   -> tests/golden/reference_run_adjoint_vectors.npz (+ a summary)

What had to be supplied from outside, because HEAD cannot run this path even on a real Paddle:
  * `functional/odeint.py` is replaced, in the adjoint module's namespace only, by `repaired_odeint` below = the
    reference's four lines (`BaseODE` -> `solver(xde, y0, rtol, atol, **options)` -> `integrate`) plus the repairs
    R1 (`xde.format` does not exist: identity), R4 (a tuple state is flattened to one fp32 vector and unpacked again
    around `func` -- the helper the reference kept for this is `utils/misc.py:1-13`), R5 (a decreasing `t_span` is
    integrated as s = -t with f~(s, y) = -f(-s, y)) and R6 (the norm callable receives the unpacked tuple);
  * `paddle.autograd.grad` of the caller's field: the field here is the fused MLP family evaluated by the oracle
    (`func(t, y)` is the CALLER'S callable in the reference), which attaches its vector-Jacobian product to the tensor
    it returns; the stand-in's `autograd.grad` calls it.  Per-trajectory products as `orc_mlp_vjp`, the batch sum of the
    parameter terms by the order-independent specification -- the one place where Paddle's own
    reduction order would be unknowable anyway.
Everything else -- which quantities enter the augmented state and in what order, the sign conventions, what is reset
between segments, the mixed / semi norm, the `grad_t_span` bookkeeping, how options reach the backward solver -- is the
reference's code executing.  The reference drops dL/dy0 (`return (None, grad_t_span, *adj_params)`, `:164-167`); it is
taken from the last augmented solve the reference runs, plus `grad_y[0]` exactly as `:157-159` adds it.

Each case stores inputs, the forward solution, `grad_y`, the parameter gradients, dL/dy0, `grad_t_span` (when asked for)
and the attempt log of every backward segment (t0 and dt in physical time: the negated solver-time values,
error ratio, accepted)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xde_oracle as xo  # noqa: E402
from oracle.ref_shim import loader  # noqa: E402
from tests.problems import fanin_weights, spiral_weights  # noqa: E402

f32 = np.float32
OUT = os.path.join(ROOT, "tests", "golden", "reference_run_adjoint_vectors.npz")


def cases():
    """(name, d, h, pre, B, t_span, solver options (forward and, by the reference's defaulting, backward), adjoint norm,
    t_requires_grad)"""
    t25 = np.linspace(0.0, 25.0, 1000).astype(f32)
    return [
        # the reference's defaults: rtol 1e-7 / atol 1e-9, mixed norm (functional/odeint_adjoint.py:175-176, 284-291)
        ("b1_cfg2_default_mixed", 2, 50, "cube", 1, t25[:6], {}, "mixed", False),
        ("b1_cfg2_seminorm", 2, 50, "cube", 1, t25[:6], {}, "seminorm", False),
        ("b1_cfg2_grad_t", 2, 50, "cube", 1, t25[:5], {}, "seminorm", True),
        ("b1_d4_options", 4, 33, "id", 1, np.linspace(0, 2, 5).astype(f32),
         dict(rtol=1e-5, atol=1e-7, first_step=0.01, max_step=0.2, safety=0.8, ifactor=5.0, dfactor=0.3), "seminorm", True),
        ("b1_d1_square", 1, 16, "square", 1, np.linspace(0, 1, 4).astype(f32), dict(rtol=1e-6, atol=1e-8), "mixed", False),
        ("b1_d3", 3, 20, "square", 1, np.linspace(0, 1, 3).astype(f32), dict(rtol=1e-6, atol=1e-8), "seminorm", False),
        ("b1_d8", 8, 48, "id", 1, np.linspace(0, 1.5, 4).astype(f32), dict(rtol=1e-6, atol=1e-8), "seminorm", True),
        ("b1_d6_rejections", 6, 24, "id", 1, np.linspace(0, 3, 4).astype(f32), dict(rtol=1e-6, atol=1e-8, wscale=3.0), "mixed", False),
        ("b1_d5_rejections_seminorm", 5, 24, "id", 1, np.linspace(0, 3, 4).astype(f32), dict(rtol=1e-6, atol=1e-8, wscale=3.0), "seminorm", False),
        ("b1_d2_reverse_span", 2, 40, "id", 1, np.linspace(1, 0, 4).astype(f32), dict(rtol=1e-6, atol=1e-8), "seminorm", True),
        ("b1_D32", 32, 64, "id", 1, np.linspace(0, 1, 3).astype(f32), dict(rtol=1e-5, atol=1e-7), "seminorm", False),
        ("b1_D64", 64, 256, "id", 1, np.linspace(0, 0.5, 3).astype(f32), dict(rtol=1e-5, atol=1e-7), "seminorm", False),
        # B > 1: one controller for the whole batch, the reference's only mode (oracle: controller="batch")
        ("batch_cfg1_default_mixed", 2, 50, "cube", 20, t25[:5], {}, "mixed", False),
        ("batch_cfg1_seminorm_grad_t", 2, 50, "cube", 20, t25[:4], {}, "seminorm", True),
        ("batch_B70_rejections", 2, 50, "id", 70, np.linspace(0, 3, 4).astype(f32), dict(rtol=1e-6, atol=1e-8, wscale=3.0), "mixed", False),
        ("batch_B45_rejections_grad_t", 2, 50, "id", 45, np.linspace(0, 3, 4).astype(f32), dict(rtol=1e-6, atol=1e-8, wscale=3.0), "mixed", True),
        ("batch_d4_seminorm_rejections", 4, 32, "id", 130, np.linspace(0, 3, 4).astype(f32), dict(rtol=1e-5, atol=1e-7, wscale=3.0), "seminorm", False),
        ("batch_d4_options", 4, 33, "id", 37, np.linspace(0, 2, 4).astype(f32),
         dict(rtol=1e-5, atol=1e-7, first_step=0.01, max_step=0.25, safety=0.8), "mixed", False),
        ("batch_d1", 1, 16, "square", 33, np.linspace(0, 1, 3).astype(f32), dict(rtol=1e-6, atol=1e-8), "seminorm", False),
    ]


def inputs(name, d, h, B, opts):
    w = spiral_weights() if (d, h) == (2, 50) and "wscale" not in opts else fanin_weights(d, h, seed=d + h)
    if "wscale" in opts:
        w = [f32(opts["wscale"]) * a for a in w]
    rng = np.random.default_rng(sum(map(ord, name)))
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    if (d, h) == (2, 50) and "wscale" not in opts:
        y0 = (np.array([2.0, 0.0]) + 0.5 * rng.standard_normal((B, 2))).astype(f32)
    return [np.ascontiguousarray(a, f32) for a in w], y0, rng


def make_repaired_odeint(ns, log):
    """functional/odeint.py:28-35 with repairs R1, R4, R5, R6 (see the module docstring)."""
    P = ns.paddle

    def repaired_odeint(func, y0, t_span, solver, *, rtol=1e-7, atol=1e-9, options={"norm": ns.ode_utils._rms_norm}):
        is_tuple = isinstance(y0, (tuple, list))
        parts = list(y0) if is_tuple else [y0]
        shapes = [list(p.shape) for p in parts]
        sizes = [int(np.prod(s)) if s else 1 for s in shapes]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        T = len(t_span)

        def unflatten(v, lead=()):
            out = tuple(v[..., int(offs[i]):int(offs[i + 1])].reshape(list(lead) + shapes[i]) for i in range(len(parts)))
            return out if is_tuple else out[0]

        def flatten(tup):
            tup = tup if is_tuple else (tup,)
            return P.concat([p.reshape([-1]) for p in tup], axis=0)

        rev = bool(t_span[1] < t_span[0])
        ts = -t_span if rev else t_span

        def flat_func(s, v):
            f = flatten(func(-s if rev else s, unflatten(v)))
            return -f if rev else f

        opts = dict(options)
        norm = opts.pop("norm")
        flat0 = flatten(y0)
        s = solver(xde=ns.BaseODE(flat_func, y0=flat0, t_span=ts), y0=flat0, rtol=rtol, atol=atol,
                   norm=lambda v: norm(unflatten(v)), **opts)
        s.func = flat_func
        if log is not None and is_tuple:  # the backward segments: record (t0, dt, ratio, accepted) from outside
            pending = {}
            orig_opt, orig_step = ns.rk.optimal_step_size, s._adaptive_step

            def opt_wrap(last_step, error_ratio, *a, **k):
                pending["dt"], pending["ratio"] = f32(last_step.a), f32(error_ratio.a)
                return orig_opt(last_step, error_ratio, *a, **k)

            def step_wrap(rk_state):
                t0 = f32(rk_state.t1.a)
                new = orig_step(rk_state)
                sgn = f32(-1.0 if rev else 1.0)  # the oracle and the kernels log physical time
                log.append((sgn * t0, sgn * pending["dt"], pending["ratio"], int(f32(new.t1.a) != t0)))
                return new

            ns.rk.optimal_step_size = opt_wrap
            s._adaptive_step = step_wrap
            try:
                sol = s.integrate(ts)
            finally:
                ns.rk.optimal_step_size = orig_opt
        else:
            sol = s.integrate(ts)
        out = unflatten(sol, lead=(T,))
        if is_tuple:
            repaired_odeint.last_tuple_solution = out
        return out

    return repaired_odeint


def run_reference(ns, om, w, y0, t, opts, adj_norm, t_grad, grad_y):
    P = ns.paddle
    params = tuple(P.to_tensor(a, dtype=P.float32) for a in w)

    def func(tt, y):
        ya = np.ascontiguousarray(y.a, f32)
        out = P.Tensor(om(0.0, ya).reshape(ya.shape))

        def vjp(cot):
            _, dy, g = om.vjp_batch(ya, np.ascontiguousarray(cot.a, f32))
            return [None, P.Tensor(dy.reshape(ya.shape))] + [P.Tensor(np.array(a)) for a in g]

        out._vjp = vjp
        return out

    log = []
    odeint_r = make_repaired_odeint(ns, log)
    ns.adjoint_mod.odeint = odeint_r
    kw = {k: v for k, v in opts.items() if k not in ("rtol", "atol", "wscale")}
    tT = P.to_tensor(t, dtype=P.float32)
    tT.stop_gradient = not t_grad
    sol = ns.adjoint_mod.odeint_adjoint(
        func, P.to_tensor(y0, dtype=P.float32), tT, rtol=opts.get("rtol", 1e-7), atol=opts.get("atol", 1e-9),
        solver=ns.Dopri5, options=dict(norm=ns.ode_utils._rms_norm, **kw),
        adjoint_options=(dict(norm="seminorm", **kw) if adj_norm == "seminorm" else None), adjoint_params=params)
    gy = P.to_tensor(grad_y, dtype=P.float32)
    res = ns.adjoint_mod.OdeintAdjointMethod.backward(sol._ctx, gy)
    assert res[0] is None and len(res) == 2 + len(params)
    grad_t = None if res[1] is None else np.ascontiguousarray(res[1].a, f32)
    gparams = [np.ascontiguousarray(r.a, f32) for r in res[2:]]
    a0 = np.ascontiguousarray(odeint_r.last_tuple_solution[2].a[1] + grad_y[0], f32)  # :153, :157-159
    return np.ascontiguousarray(sol.a, f32), gparams, a0, grad_t, np.array(log, dtype=xo.ATTEMPT_DTYPE)


def generate():
    ns = loader.load()
    out, summary = {}, []
    for name, d, h, pre, B, t, opts, adj_norm, t_grad in cases():
        w, y0, rng = inputs(name, d, h, B, opts)
        om = xo.MLP(*w, pre=pre)
        grad_y = rng.standard_normal((t.size, B, d)).astype(f32)
        sol, gparams, a0, grad_t, log = run_reference(ns, om, w, y0, t, opts, adj_norm, t_grad, grad_y)
        out[f"{name}/w1"], out[f"{name}/b1"], out[f"{name}/w2"], out[f"{name}/b2"] = w
        out[f"{name}/y0"], out[f"{name}/t"], out[f"{name}/sol"], out[f"{name}/grad_y"] = y0, t, sol, grad_y
        out[f"{name}/gw1"], out[f"{name}/gb1"], out[f"{name}/gw2"], out[f"{name}/gb2"] = gparams
        out[f"{name}/adj_y0"], out[f"{name}/log"] = a0, log
        if grad_t is not None:
            out[f"{name}/grad_t"] = grad_t
        meta = dict(pre=pre, adj_norm=adj_norm, **{k: v for k, v in opts.items() if k != "wscale"})
        out[f"{name}/meta"] = np.array(repr(meta))
        summary.append((name, sol.shape, len(log), int((log["accepted"] == 0).sum()), grad_t is not None))
    return out, summary


if __name__ == "__main__":
    xo.build()
    out, summary = generate()
    np.savez_compressed(OUT, **out)
    for name, shp, n, rej, gt in summary:
        print(f"{name:30s} solution {shp}  backward attempts {n} (rejected {rej}){'  grad_t_span' if gt else ''}")
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
