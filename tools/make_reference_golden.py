"""Golden vectors produced by the REFERENCE'S OWN CODE (unmodified files under /root/reference) running on the NumPy
`paddle` stand-in of oracle/ref_shim: adaptive Runge-Kutta solves (every tableau, solver options, step_t / jump_t,
B = 1 and B > 1 with the reference's global norm) and fixed-grid solves (Euler / Midpoint / RK4, both output
interpolants, step_size / grid_constructor grids) of the fused field family.

    python tools/make_reference_golden.py            -> tests/golden/reference_run_vectors.npz (+ a summary on stdout)

Each case stores its inputs, the reference's solution and -- for the adaptive solvers -- the reference's attempt log
(t0, dt, error ratio, accepted), captured by wrapping `optimal_step_size` / `_adaptive_step` from outside.  The vector
field is evaluated by the oracle's MLP (`func(t, y)` is the caller's callable in the reference), so what the vectors
pin is the reference's solver code.  tests/test_reference_run_golden.py compares the oracle against the file
(always) and re-derives the file from /root/reference when that tree is present (here; never on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xde_oracle as xo  # noqa: E402
from oracle.ref_shim import loader  # noqa: E402
from tests.problems import fanin_weights, spiral_weights  # noqa: E402

f32 = np.float32
OUT = os.path.join(ROOT, "tests", "golden", "reference_run_vectors.npz")
ADAPTIVE = {"Dopri5": "dopri5", "Bosh3": "bosh3", "Fehlberg2": "fehlberg2", "AdaptiveHeun": "adaptive_heun", "Dopri8": "dopri8"}


def cases():
    """(name, kind, solver, d, h, pre, B, t_span, options)"""
    t10 = np.linspace(0.0, 25.0, 1000).astype(f32)[:10]
    c = [("cfg1_dopri5_B20", "adaptive", "Dopri5", 2, 50, "cube", 20, np.linspace(0.0, 25.0, 1000).astype(f32)[:32], {}),
         ("cfg2_dopri5_B1", "adaptive", "Dopri5", 2, 50, "cube", 1, t10, {}),
         ("dopri5_B64_rejections", "adaptive", "Dopri5", 2, 50, "id", 64, np.linspace(0, 4, 9).astype(f32), dict(rtol=1e-6, atol=1e-8, wscale=3.0)),
         ("dopri5_options", "adaptive", "Dopri5", 4, 33, "id", 17, np.linspace(0, 2, 7).astype(f32),
          dict(rtol=1e-5, atol=1e-7, first_step=0.01, max_step=0.2, safety=0.8, ifactor=5.0, dfactor=0.3)),
         ("dopri5_min_step", "adaptive", "Dopri5", 3, 20, "square", 9, np.linspace(0, 1, 4).astype(f32), dict(rtol=1e-6, atol=1e-8, min_step=0.05)),
         ("dopri5_step_jump", "adaptive", "Dopri5", 2, 24, "id", 11, np.linspace(0, 1.5, 5).astype(f32),
          dict(rtol=1e-5, atol=1e-7, step_t=[0.31, 0.92, 1.2], jump_t=[0.55, 1.05])),
         ("dopri5_D64", "adaptive", "Dopri5", 64, 256, "id", 6, np.linspace(0, 1, 4).astype(f32), dict(rtol=1e-5, atol=1e-7)),
         ("bosh3", "adaptive", "Bosh3", 2, 50, "cube", 13, t10[:6], dict(rtol=1e-5, atol=1e-7)),
         ("bosh3_step_t", "adaptive", "Bosh3", 4, 32, "id", 5, np.linspace(0, 1, 4).astype(f32), dict(rtol=1e-4, atol=1e-6, step_t=[0.4, 0.77])),
         ("fehlberg2", "adaptive", "Fehlberg2", 4, 32, "id", 8, np.linspace(0, 1, 5).astype(f32), dict(rtol=1e-4, atol=1e-6)),
         ("adaptive_heun", "adaptive", "AdaptiveHeun", 1, 16, "square", 7, np.linspace(0, 1, 5).astype(f32), dict(rtol=1e-4, atol=1e-6)),
         ("dopri8", "adaptive", "Dopri8", 6, 24, "id", 10, np.linspace(0, 2, 5).astype(f32), dict(rtol=1e-7, atol=1e-9)),
         # B = 1: the global norm and the per-trajectory controller coincide -> comparable with the per-trajectory kernels
         ("b1_bosh3", "adaptive", "Bosh3", 2, 50, "cube", 1, t10[:6], dict(rtol=1e-6, atol=1e-8)),
         ("b1_fehlberg2", "adaptive", "Fehlberg2", 4, 32, "id", 1, np.linspace(0, 1, 5).astype(f32), dict(rtol=1e-4, atol=1e-6)),
         ("b1_adaptive_heun", "adaptive", "AdaptiveHeun", 1, 16, "square", 1, np.linspace(0, 1, 5).astype(f32), dict(rtol=1e-4, atol=1e-6)),
         ("b1_dopri8", "adaptive", "Dopri8", 6, 24, "id", 1, np.linspace(0, 2, 5).astype(f32), dict(rtol=1e-7, atol=1e-9)),
         ("b1_dopri5_step_jump", "adaptive", "Dopri5", 2, 24, "id", 1, np.linspace(0, 1.5, 5).astype(f32),
          dict(rtol=1e-5, atol=1e-7, step_t=[0.31, 0.92, 1.2], jump_t=[0.55, 1.05])),
         ("b1_dopri5_D64", "adaptive", "Dopri5", 64, 256, "id", 1, np.linspace(0, 1, 4).astype(f32), dict(rtol=1e-5, atol=1e-7)),
         ("b1_dopri5_D32_options", "adaptive", "Dopri5", 32, 64, "cube", 1, np.linspace(0, 1, 4).astype(f32),
          dict(rtol=1e-5, atol=1e-7, first_step=0.02, max_step=0.3, safety=0.85)),
         ("euler", "fixed", "Euler", 2, 50, "cube", 6, np.linspace(0.0, 25.0, 1000).astype(f32)[:32], {}),
         ("midpoint", "fixed", "Midpoint", 8, 48, "square", 5, np.linspace(0, 1, 9).astype(f32), {}),
         ("rk4", "fixed", "RK4", 4, 32, "id", 7, np.linspace(0, 1, 9).astype(f32), {}),
         ("rk4_cubic", "fixed", "RK4", 2, 50, "cube", 4, np.linspace(0, 1, 9).astype(f32), dict(interp="cubic")),
         ("rk4_step_size", "fixed", "RK4", 2, 50, "cube", 9, np.linspace(0, 1, 6).astype(f32), dict(step_size=0.05)),
         ("euler_step_size", "fixed", "Euler", 32, 64, "id", 5, np.linspace(0, 1, 6).astype(f32), dict(step_size=0.07)),
         ("midpoint_grid_constructor", "fixed", "Midpoint", 4, 32, "id", 6, np.linspace(0, 1, 5).astype(f32),
          dict(grid=[0.0, 0.2, 0.3, 0.55, 0.8, 0.9, 1.0]))]
    return c


def inputs(name, d, h, pre, B, opts):
    w = spiral_weights() if (d, h) == (2, 50) and "wscale" not in opts else fanin_weights(d, h, seed=d + h)
    if "wscale" in opts:
        w = [f32(opts["wscale"]) * a for a in w]
    rng = np.random.default_rng(abs(hash(name)) % (1 << 31) if False else sum(map(ord, name)))
    y0 = rng.uniform(-1, 1, (B, d)).astype(f32)
    if (d, h) == (2, 50) and "wscale" not in opts:
        y0 = (np.array([2.0, 0.0]) + 0.5 * rng.standard_normal((B, 2))).astype(f32)
    return [np.ascontiguousarray(a, f32) for a in w], y0


def run_reference(ns, kind, solver, om, y0, t, opts):
    """The reference's solver classes on the stand-in.  -> (solution, log | None)"""
    P = ns.paddle
    func = lambda tt, y: P.Tensor(om(0.0, np.ascontiguousarray(y.a, f32)).reshape(y.a.shape))  # noqa: E731
    tT = P.to_tensor(t, dtype=P.float32)
    log = []
    if kind == "adaptive":
        y0T = P.to_tensor(y0, dtype=P.float32)
        kw = {k: v for k, v in opts.items() if k not in ("rtol", "atol", "wscale")}
        s = getattr(ns, solver)(xde=ns.BaseODE(func, y0=y0T, t_span=tT), y0=y0T, rtol=opts.get("rtol", 1e-7),
                                atol=opts.get("atol", 1e-9), norm=ns.ode_utils._rms_norm, **kw)
        s.func = func  # AdaptiveRKSolver._adaptive_step reads self.func after a jump (:271); odeint never sets it
        pending = {}
        orig_opt, orig_step = ns.rk.optimal_step_size, s._adaptive_step

        def opt_wrap(last_step, error_ratio, *a, **k):
            pending["dt"], pending["ratio"] = f32(last_step.a), f32(error_ratio.a)
            return orig_opt(last_step, error_ratio, *a, **k)

        def step_wrap(rk_state):
            t0 = f32(rk_state.t1.a)
            new = orig_step(rk_state)
            log.append((t0, pending["dt"], pending["ratio"], int(f32(new.t1.a) != t0)))
            return new

        ns.rk.optimal_step_size = opt_wrap
        s._adaptive_step = step_wrap
        try:
            sol = s.integrate(tT).a
        finally:
            ns.rk.optimal_step_size = orig_opt
        return np.ascontiguousarray(sol, f32), np.array(log, dtype=xo.ATTEMPT_DTYPE)
    y0T = P.to_tensor(y0[:, None, :], dtype=P.float32)  # [B, 1, D]: the layout the fixed solvers document
    kw = {}
    if "step_size" in opts:
        kw["step_size"] = opts["step_size"]
    if "grid" in opts:
        g = np.asarray(opts["grid"], f32)
        kw["grid_constructor"] = lambda y, tt: P.to_tensor(g, dtype=P.float32)
    s = getattr(ns, solver)(xde=ns.BaseODE(func, y0=y0T, t_span=tT), y0=y0T, interp=opts.get("interp", "linear"),
                            rtol=1e-7, atol=1e-9, norm=ns.ode_utils._rms_norm, **kw)
    return np.ascontiguousarray(s.integrate(tT).a, f32), None


def generate():
    ns = loader.load()
    out, summary = {}, []
    for name, kind, solver, d, h, pre, B, t, opts in cases():
        w, y0 = inputs(name, d, h, pre, B, opts)
        om = xo.MLP(*w, pre=pre)
        sol, log = run_reference(ns, kind, solver, om, y0, t, opts)
        out[f"{name}/w1"], out[f"{name}/b1"], out[f"{name}/w2"], out[f"{name}/b2"] = w
        out[f"{name}/y0"], out[f"{name}/t"], out[f"{name}/sol"] = y0, t, sol
        if log is not None:
            out[f"{name}/log"] = log
        meta = dict(kind=kind, solver=solver, pre=pre, **{k: v for k, v in opts.items() if k != "wscale"})
        out[f"{name}/meta"] = np.array(repr(meta))
        summary.append((name, sol.shape, None if log is None else (len(log), int((log["accepted"] == 0).sum()))))
    return out, summary


if __name__ == "__main__":
    xo.build()
    out, summary = generate()
    np.savez_compressed(OUT, **out)
    for name, shp, lg in summary:
        print(f"{name:28s} solution {shp}" + ("" if lg is None else f"  attempts {lg[0]} (rejected {lg[1]})"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
