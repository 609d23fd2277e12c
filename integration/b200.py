"""Reference-side binding of libxde_b200.so -- the ONE file a PaddleXDE maintainer adds (as `paddlexde/solver/b200.py`).

It needs nothing from this repository except the shared library and `include/xde_b200.h`'s layouts: `ctypes` + Paddle.
`odeint(func, y0, t, solver=Dopri5B200)` keeps the reference's call site (`functional/odeint.py:30-31` instantiates
`solver(xde=xde, y0=xde.y0, rtol=rtol, atol=atol, **options)` and calls `.integrate(t_span)`); `OdeintAdjointB200` is the
twin of `OdeintAdjointMethod` (`functional/odeint_adjoint.py:11-167`).  INTEGRATION.md walks through it.

Executed in this repository's CPU suite (tests/test_integration_stub.py): the reference's own `odeint` drives these
classes on the NumPy `paddle` stand-in, with the C ABI answered by the CPU oracle -- argument order, struct layouts and
the protocol are checked against reference-run vectors there; on a B200 the same file binds the real library."""
import ctypes as C
import os

import numpy as np
import paddle

_lib = C.CDLL(os.environ.get("XDE_B200_LIBRARY", "libxde_b200.so"))          # include/xde_b200.h

XDE_CTRL_TRAJECTORY, XDE_CTRL_BATCH = 0, 1
XDE_ADJ_NORM_MIXED, XDE_ADJ_NORM_SEMI = 0, 1
XDE_PRE = {"id": 0, "square": 1, "cube": 2}
_STATUS = {1: "underflow in dt", 2: "non-finite values in state `y`", 3: "max_num_steps exceeded"}


class _Field(C.Structure):               # xde_mlp_field_t
    _fields_ = [("d", C.c_int32), ("h", C.c_int32), ("pre", C.c_int32), ("_pad", C.c_int32),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


class _Opts(C.Structure):                # xde_ctrl_opts_t  (AdaptiveRKSolver.__init__ kwargs, base_adaptive_solver_rk.py:32-49)
    _fields_ = [("rtol", C.c_float), ("atol", C.c_float), ("min_step", C.c_float), ("max_step", C.c_float),
                ("first_step", C.c_float), ("safety", C.c_float), ("ifactor", C.c_float), ("dfactor", C.c_float),
                ("max_num_steps", C.c_int32), ("_pad", C.c_int32)]


class _Stats(C.Structure):               # xde_stats_t
    _fields_ = [("n_attempts", C.c_ulonglong), ("n_accepted", C.c_ulonglong), ("nfe", C.c_ulonglong),
                ("status", C.c_int32), ("_pad", C.c_int32)]


_lib.xde_dopri5_mlp_f32.restype = C.c_int
_lib.xde_dopri5_mlp_f32.argtypes = [C.POINTER(_Field), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                    C.POINTER(_Opts), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.xde_dopri5_mlp_adjoint_f32.restype = C.c_int
_lib.xde_dopri5_mlp_adjoint_f32.argtypes = [C.POINTER(_Field), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.POINTER(_Opts), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
_lib.xde_last_error.restype = C.c_char_p


def _stream():
    return paddle.device.cuda.current_stream().cuda_stream


def _field_of(func, pre="cube"):
    """example/ode_demo.py:17-33: func.net = Sequential(Linear(2,50), Tanh(), Linear(50,2)) applied to y**3.
    Paddle nn.Linear stores weight [in, out] -- exactly the layout xde_mlp_field_t expects."""
    l1, act, l2 = list(func.net.children())
    assert type(act).__name__ == "Tanh", "only tanh MLP fields are fused; there is no fallback"
    keep = [p.astype("float32").contiguous() for p in (l1.weight, l1.bias, l2.weight, l2.bias)]
    f = _Field(l1.weight.shape[0], l1.weight.shape[1], XDE_PRE[getattr(func, "pre", pre)], 0,
               *[p.data_ptr() for p in keep])
    return f, keep


def _make_opts(rtol, atol, min_step=0.0, max_step=float("inf"), first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
               max_num_steps=2 ** 31 - 1):
    return _Opts(rtol, atol, min_step, max_step, float("nan") if first_step is None else first_step, safety, ifactor,
                 dfactor, max_num_steps, 0)


def _check(rc, stats):
    if rc != 0:
        raise RuntimeError(_lib.xde_last_error().decode())
    st = _Stats.from_buffer_copy(stats.numpy().tobytes())       # synchronises, like the reference's asserts
    assert st.status == 0, _STATUS.get(st.status, f"solver status {st.status}")
    return st


class Dopri5B200:
    """solver=Dopri5B200: replaces AdaptiveSolver.integrate (solver/base_adaptive_solver.py:24-31) and everything under it
    (base_adaptive_solver_rk.py:116-292, utils/ode_utils.py:28-97) by one kernel launch."""
    order = 5

    def __init__(self, xde, y0, rtol, atol, min_step=0.0, max_step=float("inf"), first_step=None, safety=0.9,
                 ifactor=10.0, dfactor=0.2, max_num_steps=2 ** 31 - 1, norm=None, controller=XDE_CTRL_TRAJECTORY, **unused):
        self.xde, self.y0, self.controller = xde, y0, controller
        self.opts = _make_opts(rtol, atol, min_step, max_step, first_step, safety, ifactor, dfactor, max_num_steps)

    def integrate(self, t_span):
        field, keep = _field_of(self.xde.func)
        y0 = self.y0.astype("float32").contiguous()
        t = t_span.astype("float32").contiguous()     # device tensor; must be strictly monotone
        B, T = int(np.prod(y0.shape[:-1])), t.shape[0]
        out = paddle.empty([T] + list(y0.shape), dtype="float32")
        stats = paddle.zeros([4], dtype="int64")      # 32 bytes = sizeof(xde_stats_t)
        rc = _lib.xde_dopri5_mlp_f32(C.byref(field), y0.data_ptr(), B, t.data_ptr(), T, C.byref(self.opts),
                                     self.controller, out.data_ptr(), stats.data_ptr(), None, _stream())
        _check(rc, stats)
        return out                                    # [T, *y0.shape], as base_adaptive_solver.py:25


class OdeintAdjointB200(paddle.autograd.PyLayer):
    """Twin of OdeintAdjointMethod (functional/odeint_adjoint.py:11-167): forward = the fused solve, backward = one launch
    of the augmented reverse-time solve.  holder: {"func", "odeint", "rtol", "atol", "options": {...solver kwargs}}."""

    @staticmethod
    def forward(ctx, holder, y0, t_span, *params):
        with paddle.no_grad():
            ans = holder["odeint"](holder["func"], y0, t_span, solver=Dopri5B200, rtol=holder["rtol"], atol=holder["atol"],
                                   options=holder.get("options", {}))
        ctx.holder, ctx.t_requires_grad = holder, not t_span.stop_gradient
        ctx.save_for_backward(t_span, ans)
        return ans

    @staticmethod
    def backward(ctx, grad_y):
        t_span, ans = ctx.saved_tensor()
        h = ctx.holder
        field, keep = _field_of(h["func"])
        opts = _make_opts(h["rtol"], h["atol"], **{k: v for k, v in h.get("options", {}).items()
                                                   if k in ("min_step", "max_step", "first_step", "safety", "ifactor",
                                                            "dfactor", "max_num_steps")})
        T, B = t_span.shape[0], int(np.prod(ans.shape[1:-1]))
        t = t_span.astype("float32").contiguous()
        gy = grad_y.astype("float32").contiguous()
        g = paddle.zeros([field.d * field.h * 2 + field.h + field.d], dtype="float32")      # (gW1, gb1, gW2, gb2)
        grad_t = paddle.empty([T], dtype="float32") if ctx.t_requires_grad else None          # :129-141,161-162
        stats = paddle.zeros([4], dtype="int64")
        rc = _lib.xde_dopri5_mlp_adjoint_f32(C.byref(field), t.data_ptr(), T, ans.data_ptr(), gy.data_ptr(), B,
                                             C.byref(opts), XDE_CTRL_TRAJECTORY, XDE_ADJ_NORM_SEMI,
                                             g.data_ptr(), None, grad_t.data_ptr() if grad_t is not None else None,
                                             stats.data_ptr(), None, _stream())
        _check(rc, stats)
        d, hh = field.d, field.h
        gw1, gb1, gw2, gb2 = paddle.split(g, [d * hh, hh, hh * d, d])
        return (None, grad_t, gw1.reshape([d, hh]), gb1, gw2.reshape([hh, d]), gb2)    # y0: None (:167)
