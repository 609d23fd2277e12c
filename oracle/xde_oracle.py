"""ctypes binding of the C oracle (oracle/libxde_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package.  See xde_oracle.h for the parity statement and citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libxde_oracle.so")

PRE = {"id": 0, "identity": 0, "square": 1, "cube": 2}
CTRL = {"trajectory": 0, "batch": 1}
ADJ_NORM = {"mixed": 0, "default": 0, "seminorm": 1}
FIXED = {"euler": 0, "rk4": 1, "midpoint": 2}
RK = {"dopri5": 0, "bosh3": 1, "fehlberg2": 2, "adaptive_heun": 3, "dopri8": 4}
SDE = {"em": 0, "euler": 0, "milstein": 1}
INTERP = {"linear": 0, "cubic": 1, "hermite": 1, "bez": 2, "bezier": 2}
STATUS = {0: "OK", 1: "DT_UNDERFLOW", 2: "NONFINITE_STATE", 3: "MAX_STEPS", 4: "BAD_ARG", 5: "INTERP_RANGE"}


class _Mlp(C.Structure):
    _fields_ = [("d", C.c_int32), ("h", C.c_int32), ("pre", C.c_int32),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


class _Opts(C.Structure):
    _fields_ = [("rtol", C.c_float), ("atol", C.c_float), ("min_step", C.c_float), ("max_step", C.c_float),
                ("first_step", C.c_float), ("safety", C.c_float), ("ifactor", C.c_float),
                ("dfactor", C.c_float), ("max_num_steps", C.c_int32)]


STATS_DTYPE = np.dtype([("n_attempts", np.int64), ("n_accepted", np.int64), ("nfe", np.int64),
                        ("status", np.int32), ("min_abs_ratio_m1", np.float32)], align=True)
ATTEMPT_DTYPE = np.dtype([("t0", np.float32), ("dt", np.float32), ("ratio", np.float32),
                          ("accepted", np.int32)], align=True)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only; seconds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "xde_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_tanhf.restype = C.c_float
        _lib.orc_tanhf.argtypes = [C.c_float]
        _lib.orc_root5f.restype = C.c_float
        _lib.orc_root5f.argtypes = [C.c_float]
        _lib.orc_rootpf.restype = C.c_float
        _lib.orc_rootpf.argtypes = [C.c_float, C.c_int32]
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class MLP:
    """tanh(pre(y) @ w1 + b1) @ w2 + b2 with Paddle nn.Linear weight layout [in, out]."""
    w1: np.ndarray
    b1: np.ndarray
    w2: np.ndarray
    b2: np.ndarray
    pre: str = "cube"

    def __post_init__(self):
        self.w1, self.b1, self.w2, self.b2 = (_f32(a) for a in (self.w1, self.b1, self.w2, self.b2))
        self.d, self.h = self.w1.shape
        assert self.w2.shape == (self.h, self.d) and self.b1.shape == (self.h,) and self.b2.shape == (self.d,)

    @property
    def n_params(self):
        return 2 * self.d * self.h + self.h + self.d

    def c(self):
        return _Mlp(self.d, self.h, PRE[self.pre], _p(self.w1), _p(self.b1), _p(self.w2), _p(self.b2))

    def split(self, flat):
        d, h = self.d, self.h
        o = 0
        out = []
        for shp in ((d, h), (h,), (h, d), (d,)):
            n = int(np.prod(shp))
            out.append(np.asarray(flat[o:o + n]).reshape(shp))
            o += n
        return out

    # generic-callable views used with oracle_np
    def __call__(self, t, y):
        y = _f32(y)
        f = np.empty_like(y)
        m = self.c()
        lib().orc_mlp_eval_batch(C.byref(m), _p(y), C.c_int64(y.size // self.d), _p(f))
        return f

    def vjp(self, t, y, c):
        y, c = _f32(y), _f32(c)
        B = y.size // self.d
        f, dy = np.empty_like(y), np.empty_like(y)
        g = np.zeros(self.n_params, np.float32)
        gw1, gb1, gw2, gb2 = self.split(g)
        m = self.c()
        y2, c2, f2, dy2 = (a.reshape(B, self.d) for a in (y, c, f, dy))
        for b in range(B):
            lib().orc_mlp_vjp(C.byref(m), _p(y2[b]), _p(c2[b]), _p(f2[b]), _p(dy2[b]),
                              _p(gw1), _p(gb1), _p(gw2), _p(gb2))
        return f, dy, [gw1, gb1, gw2, gb2]

    def vjp_batch(self, y, c):
        """paddle.autograd.grad(f(y), (y, *params), grad_outputs=c) over a batch: y, c [B, D] -> (f, dy, [gW1, gb1, gW2,
        gb2]) with the parameter gradients summed over the batch by the order-independent specification."""
        y, c = _f32(y).reshape(-1, self.d), _f32(c).reshape(-1, self.d)
        f, dy = np.empty_like(y), np.empty_like(y)
        g = np.zeros(self.n_params, np.float32)
        m = self.c()
        lib().orc_mlp_vjp_batch(C.byref(m), _p(y), _p(c), C.c_int64(y.shape[0]), _p(f), _p(dy), _p(g))
        return f, dy, self.split(g)


def make_opts(rtol=1e-7, atol=1e-9, min_step=0.0, max_step=float("inf"), first_step=None, safety=0.9,
              ifactor=10.0, dfactor=0.2, max_num_steps=2**31 - 1):
    return _Opts(rtol, atol, min_step, max_step, float("nan") if first_step is None else first_step,
                 safety, ifactor, dfactor, max_num_steps)


def tanhf(x):
    return np.array([lib().orc_tanhf(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32).reshape(np.shape(x))


def root5f(x):
    return np.array([lib().orc_root5f(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32).reshape(np.shape(x))


def rootpf(x, p):
    return np.array([lib().orc_rootpf(float(v), int(p)) for v in np.asarray(x, np.float32).ravel()],
                    np.float32).reshape(np.shape(x))


def dopri5_mlp(mlp: MLP, y0, t_span, **kw):
    """-> (out [T,B,D], stats recarray, log recarray|None, status)"""
    return adaptive_rk_mlp("dopri5", mlp, y0, t_span, **kw)


def sort_tvals(tvals, t_span):
    """sort_tvals (utils/ode_utils.py:22-25) in solver time: keep values >= t_span[0], ascending; a decreasing
    t_span is integrated as s = -t (repair R5)."""
    if tvals is None:
        return np.zeros(0, np.float32)
    v = _f32(tvals).reshape(-1)
    t0 = np.float32(t_span[0])
    if t_span[1] < t_span[0]:
        v, t0 = -v, -t0
    return np.sort(v[v >= t0]).astype(np.float32)


def adaptive_rk_mlp(method: str, mlp: MLP, y0, t_span, *, controller="trajectory", log_traj: Optional[int] = None,
                    log_cap=100000, nthreads=0, step_t=None, jump_t=None, **opt_kw):
    """Any embedded tableau of the reference (RK keys) -> (out [T,B,D], stats, log|None, status)"""
    y0, t_span = _f32(y0), _f32(t_span)
    B, D = y0.shape
    T = t_span.size
    out = np.empty((T, B, D), np.float32)
    ns = B if controller == "trajectory" else 1
    stats = np.zeros(ns, STATS_DTYPE)
    want_log = log_traj is not None or controller == "batch"
    log = np.zeros(log_cap if want_log else 0, ATTEMPT_DTYPE)
    log_len = C.c_int64(0)
    m, o = mlp.c(), make_opts(**opt_kw)
    st_, jt_ = sort_tvals(step_t, t_span), sort_tvals(jump_t, t_span)
    rc = lib().orc_adaptive_rk_mlp_grid(C.c_int32(RK[method]), C.byref(m), _p(y0), C.c_int64(B), _p(t_span),
                                        C.c_int32(T), C.byref(o), C.c_int32(CTRL[controller]),
                                        _p(st_) if st_.size else None, C.c_int32(st_.size),
                                        _p(jt_) if jt_.size else None, C.c_int32(jt_.size), _p(out), _p(stats),
                                        _p(log) if want_log else None, C.c_int64(log_cap),
                                        C.c_int64(log_traj or 0), C.byref(log_len), C.c_int32(nthreads))
    return out, stats.view(np.recarray), (log[:log_len.value].view(np.recarray) if want_log else None), rc


def fixed_mlp(method: str, mlp: MLP, y0, t_span, nthreads=0):
    """-> out [B,T,D]"""
    y0, t_span = _f32(y0), _f32(t_span)
    B, D = y0.shape
    out = np.empty((B, t_span.size, D), np.float32)
    m = mlp.c()
    rc = lib().orc_fixed_mlp(C.c_int32(FIXED[method]), C.byref(m), _p(y0), C.c_int64(B), _p(t_span),
                             C.c_int32(t_span.size), _p(out), C.c_int32(nthreads))
    assert rc == 0, STATUS.get(rc, rc)
    return out


def dopri5_mlp_adjoint(mlp: MLP, t_span, y_ans, grad_y, *, controller="trajectory", adj_norm="seminorm",
                       log_traj: Optional[int] = None, log_cap=100000, nthreads=0, grad_t=None, **opt_kw):
    """-> (gparams flat [P], adj_y0 [B,D], stats, log|None, status); grad_t: optional float32 [T] array that
    receives grad_t_span (the reference's t_requires_grad branch)"""
    t_span, y_ans, grad_y = _f32(t_span), _f32(y_ans), _f32(grad_y)
    T, B, D = y_ans.shape
    g = np.zeros(mlp.n_params, np.float32)
    a0 = np.zeros((B, D), np.float32)
    ns = B if controller == "trajectory" else 1
    stats = np.zeros(ns, STATS_DTYPE)
    want_log = log_traj is not None or controller == "batch"
    log = np.zeros(log_cap if want_log else 0, ATTEMPT_DTYPE)
    log_len = C.c_int64(0)
    m, o = mlp.c(), make_opts(**opt_kw)
    rc = lib().orc_dopri5_mlp_adjoint(C.byref(m), _p(t_span), C.c_int32(T), _p(y_ans), _p(grad_y),
                                      C.c_int64(B), C.byref(o), C.c_int32(CTRL[controller]),
                                      C.c_int32(ADJ_NORM[adj_norm]), _p(g), _p(a0), _p(grad_t), _p(stats),
                                      _p(log) if want_log else None, C.c_int64(log_cap),
                                      C.c_int64(log_traj or 0), C.byref(log_len), C.c_int32(nthreads))
    return g, a0, stats.view(np.recarray), (log[:log_len.value].view(np.recarray) if want_log else None), rc


def sde_mlp(scheme: str, drift: MLP, diffusion: MLP, y0, t_span, dW, nthreads=0):
    """-> out [B,T,D]; dW [T-1,B,D]"""
    y0, t_span, dW = _f32(y0), _f32(t_span), _f32(dW)
    B, D = y0.shape
    T = t_span.size
    assert dW.shape == (T - 1, B, D)
    out = np.empty((B, T, D), np.float32)
    f, g = drift.c(), diffusion.c()
    rc = lib().orc_sde_mlp(C.c_int32(SDE[scheme]), C.byref(f), C.byref(g), _p(y0), C.c_int64(B), _p(t_span),
                           C.c_int32(T), _p(dW), _p(out), C.c_int32(nthreads))
    assert rc == 0, STATUS.get(rc, rc)
    return out


def sde_mlp_adjoint(drift: MLP, diffusion: MLP, t_span, y_all, grad_y, dW, nthreads=0):
    """Exact adjoint of the Euler-Maruyama recursion -> (g_drift flat, g_diffusion flat, adj_y0 [B,D]);
    y_all, grad_y [B,T,D], dW [T-1,B,D]."""
    t_span, y_all, grad_y, dW = _f32(t_span), _f32(y_all), _f32(grad_y), _f32(dW)
    B, T, D = y_all.shape
    assert dW.shape == (T - 1, B, D) and grad_y.shape == y_all.shape
    gf, gg = np.zeros(drift.n_params, np.float32), np.zeros(diffusion.n_params, np.float32)
    a0 = np.zeros((B, D), np.float32)
    f, g = drift.c(), diffusion.c()
    rc = lib().orc_sde_mlp_adjoint(C.byref(f), C.byref(g), _p(t_span), C.c_int32(T), _p(y_all), _p(grad_y),
                                   C.c_int64(B), _p(dW), _p(gf), _p(gg), _p(a0), C.c_int32(nthreads))
    assert rc == 0, STATUS.get(rc, rc)
    return gf, gg, a0


def history_gather(kind: str, his, his_span, lags):
    """his [..., Th, D] -> (values, derivs) [..., L, D]"""
    his, his_span, lags = _f32(his), _f32(his_span), _f32(np.atleast_1d(lags))
    lead, Th, D = his.shape[:-2], his.shape[-2], his.shape[-1]
    R = int(np.prod(lead)) if lead else 1
    L = lags.size
    val = np.empty((R, L, D), np.float32)
    der = np.empty((R, L, D), np.float32)
    rc = lib().orc_history_gather(C.c_int32(INTERP[kind]), _p(his), C.c_int64(R), C.c_int32(Th), C.c_int32(D),
                                  _p(his_span), _p(lags), C.c_int32(L), _p(val), _p(der))
    assert rc == 0, STATUS.get(rc, rc)
    return val.reshape(lead + (L, D)), der.reshape(lead + (L, D))


def history_gather_bwd(grad_y, deriv):
    grad_y, deriv = _f32(grad_y), _f32(deriv)
    L, D = grad_y.shape[-2:]
    R = grad_y.size // (L * D)
    g = np.empty(L, np.float32)
    lib().orc_history_gather_bwd(_p(grad_y), _p(deriv), C.c_int64(R), C.c_int32(L), C.c_int32(D), _p(g))
    return g


def dde_fuse(dy, dt, y0):
    dy, y0 = _f32(dy), _f32(y0)
    out = np.empty_like(y0)
    lib().orc_dde_fuse(_p(dy), C.c_float(dt), _p(y0), C.c_int64(y0.size), _p(out))
    return out
