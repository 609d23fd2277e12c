"""Euler / RK4 (3/8 rule) with the reference's FixedSolver protocol
(paddlexde/solver/base_fixed_solver.py:17-64 constructor, :103-144 integrate; fixed_solver/euler.py,
fixed_solver/rk4.py).  `integrate` is one kernel launch over the whole grid."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _tensor as T
from .._lib import FIXED, SDE, ST_TC_RANGE, XDE_E_UNSUPPORTED_FIELD, UnsupportedFieldError, check, lib
from .adaptive_solver import device_tspan, host_tspan


class FixedSolver:
    order: int
    method: str

    def __init__(self, xde, y0, step_size=None, grid_constructor=None, interp="linear", perturb=False,
                 out_stride=1, math="auto", check_status=True, **kwargs):
        if step_size is not None and grid_constructor is not None:
            raise ValueError("step_size and grid_constructor are mutually exclusive arguments.")
        # step_size / grid_constructor (base_fixed_solver.py:49-89).  The reference's loop (:119-139) takes its first
        # len(t_span) steps on the constructed grid and reports, for output i, the LINEAR interpolant of step i evaluated
        # at t_span[i] -- an extrapolation when the grid is finer than t_span.  Reproduced as it is: the fused kernel
        # integrates over grid[:len(t_span)], a second kernel applies linear_interp (interp_fn.py:4-10).
        self.step_size, self.grid_constructor = step_size, grid_constructor
        if (step_size is not None or grid_constructor is not None) and interp == "cubic":
            raise NotImplementedError("interp='cubic' on a step_size / grid_constructor grid needs f at the grid points "
                                      "(cubic_hermite_interp, interp_fn.py:13-20); the fused path offers interp='linear'")
        # Output interpolation (base_fixed_solver.py:133-139).  The grid IS t_span, so every output time is the
        # end of its step: linear_interp returns y1 (interp_fn.py:7-8) and cubic_hermite_interp evaluates at
        # h = 1, i.e. h00 = h10 = h11 = 0, h01 = 1 -> y1 as well (interp_fn.py:13-20).  Both run the same kernel.
        if interp not in ("linear", "cubic", "", None):
            raise ValueError(f"interp must be 'linear' or 'cubic', got {interp!r}")
        for key in ("atol", "rtol"):  # base_fixed_solver.py:45-47 requires them
            if key not in kwargs:
                raise KeyError(key)
        self.xde, self.y0 = xde, y0
        self.out_stride = int(out_stride)
        # math="tensor": tcgen05 path (csrc/xde_tc.cu) for D in {16,32,64}: fp16-split 3-product GEMMs with
        #   fp32 accumulation, as accurate against fp64 as the FP32 kernels, not bit-identical to them;
        # math="fp32": FFMA kernels, bit-exact against the oracle's arithmetic specification;
        # math="auto" (default): the tensor path where a kernel exists for the shape, the FP32 kernels
        #   otherwise (small states are always FP32: K = 2 is degenerate for an MMA).
        if math not in ("auto", "fp32", "tensor"):
            raise ValueError(f"math must be 'auto', 'fp32' or 'tensor', got {math!r}")
        self.math = math
        # the tensor-core kernels report stage inputs outside their fp16 operand range (|pre(y)| >= 65504 or
        # non-finite) through a device status word; reading it synchronises (like the reference's asserts do).
        # check_status=False stays asynchronous: the caller vouches for the range.
        self.check_status = check_status

    def _launch(self, tensor_call, fp32_call, dev):
        """Both are CUDA kernels of libxde_b200.  "auto" asks the tensor-core one first and takes the FP32 one
        when it has no kernel for the shape (XDE_E_UNSUPPORTED_FIELD) or reports XDE_ST_TC_RANGE; "tensor"
        raises in those cases (no silent change of arithmetic); "fp32" never touches the tensor cores."""
        if self.math == "fp32":
            return check(fp32_call())
        status = T.zeros((1,), dev, "i32") if self.check_status else None
        rc = tensor_call(T.ptr(status))
        if rc == XDE_E_UNSUPPORTED_FIELD and self.math == "auto":
            return check(fp32_call())
        check(rc)
        if status is not None and int(T.to_host(status)[0]) == ST_TC_RANGE:
            if self.math == "auto":
                return check(fp32_call())
            raise OverflowError("a stage input left the fp16 operand range of the tensor-core kernels "
                                "(|pre(y)| >= 65504 or non-finite): use math='fp32' (or 'auto')")

    def _time_grid(self, t_host):
        """grid_constructor(y0, t) / _grid_constructor_from_step_size (base_fixed_solver.py:66-89) in fp32."""
        if self.grid_constructor is not None:
            g = np.asarray(T.to_host(self.grid_constructor(self.y0, t_host)), dtype=np.float32).reshape(-1)
        else:
            start, end = t_host[0], t_host[-1]
            niters = int(np.ceil(np.float32((end - start) / np.float32(self.step_size)) + np.float32(1.0)))
            g = np.arange(0, niters, dtype=np.float32) * np.float32(self.step_size) + start
            g[-1] = end
        if g.size < 1 or g[0] != t_host[0] or g[-1] != t_host[-1]:  # the reference's two asserts (:116-117)
            raise AssertionError("the time grid must start at t_span[0] and end at t_span[-1]")
        if g.size < t_host.size:
            raise ValueError(f"the time grid has {g.size} points but the solver takes len(t_span) - 1 = {t_host.size - 1} steps "
                             "on it (base_fixed_solver.py:119-121)")
        return np.ascontiguousarray(g[:t_host.size])

    def integrate(self, t_span):
        if self.step_size is not None or self.grid_constructor is not None:
            return self._integrate_on_grid(t_span)
        return self._integrate(t_span)

    def _integrate_on_grid(self, t_span):
        if self.out_stride != 1:
            raise ValueError("out_stride needs grid == t_span")
        t_host = host_tspan(t_span)
        grid = self._time_grid(t_host)
        host_tspan(grid)  # strictly monotone, like every grid the kernels integrate over
        y0 = T.to_dev(self.y0)
        D = y0.shape[-1]
        B = y0.numel() // D
        Tn = t_host.size
        # the grid solution in the kernels' own [B, T, D] layout
        saved = self.y0
        try:
            self.y0 = y0.reshape(B, 1, D)
            y_grid = T.to_dev(self._integrate(grid), like=y0)
        finally:
            self.y0 = saved
        out = T.empty((B, Tn, D), y0)
        check(lib().xde_fixed_interp_linear_f32(T.ptr(y_grid), T.ptr(device_tspan(grid, y0)), T.ptr(device_tspan(t_host, y0)),
                                                B, Tn, D, T.ptr(out), T.stream(y0)))
        return self._layout(out, y0, B, Tn, D)

    def _integrate(self, t_span):
        kind = getattr(self.xde, "kind", None)
        y0 = T.to_dev(self.y0)
        t_host = host_tspan(t_span)
        t_dev = device_tspan(t_host, y0)
        Tn = t_host.size
        D = y0.shape[-1]
        B = y0.numel() // D
        n_out = (Tn - 1 + self.out_stride - 1) // self.out_stride + 1
        out = T.empty((B, n_out, D), y0)
        stream = T.stream(y0)
        # the kernels are specialised on the field's state dim and index y0 / out / dW with it: a mismatch would read
        # and write out of bounds on the device (the C ABI sees only pointers), so it is refused here
        fields = {"ode": ("field",), "sde": ("drift", "diffusion")}.get(kind, ())
        for name in fields:
            fd = getattr(self.xde, name).d
            if D != fd:
                raise ValueError(f"y0 last dim {D} != state dim {fd} of the {name} field")
        if kind == "ode":
            fs = self.xde.field.c_struct()
            args = (FIXED[self.method], C.byref(fs), T.ptr(y0), B, T.ptr(t_dev), Tn, self.out_stride, T.ptr(out))
            self._launch(lambda st: lib().xde_rk_fixed_mlp_tc_f32(*args, st, stream),
                         lambda: lib().xde_rk_fixed_mlp_f32(*args, stream), y0)
        elif kind == "sde":
            if self.method != "euler" and self.xde.scheme == "em":
                raise UnsupportedFieldError("sdeint is fused for solver=Euler (Euler-Maruyama) only")
            f, g = self.xde.drift.c_struct(), self.xde.diffusion.c_struct()
            if self.xde.bm_seed is not None:  # device-side Philox increments: no [T-1, B, D] table at all
                def philox(math, st=None):
                    return lib().xde_sde_mlp_philox_f32(SDE[self.xde.scheme], math, C.byref(f), C.byref(g), T.ptr(y0), B,
                                                        T.ptr(t_dev), Tn, int(self.xde.bm_seed) & (2 ** 64 - 1),
                                                        self.xde.bm_offset, self.out_stride, T.ptr(out), st, stream)
                self._launch(lambda st: philox(1, st), lambda: philox(0), y0)
            else:
                dW = T.to_dev(self.xde.bm_increments, like=y0)
                if tuple(dW.shape) != (Tn - 1, B, D):
                    raise ValueError(f"bm_increments must be [T-1, B, D] = {(Tn - 1, B, D)}, got {tuple(dW.shape)}")
                args = (SDE[self.xde.scheme], C.byref(f), C.byref(g), T.ptr(y0), B, T.ptr(t_dev), Tn, T.ptr(dW),
                        self.out_stride, T.ptr(out))
                self._launch(lambda st: lib().xde_sde_mlp_tc_f32(*args, st, stream),
                             lambda: lib().xde_sde_mlp_f32(*args, stream), y0)
        else:
            raise UnsupportedFieldError(f"fixed solvers integrate ODE/SDE problems on the device, not {kind!r}")
        return self._layout(out, y0, B, n_out, D)

    def _layout(self, out, y0, B, n_out, D):
        # concat(axis=-2) of the per-time states (base_fixed_solver.py:143)
        shp = tuple(y0.shape)
        if len(shp) >= 3 and shp[-2] == 1:          # y0 [..., 1, D] -> [..., T, D]
            res = out.reshape(shp[:-2] + (n_out, D))
        elif len(shp) == 1:
            res = out.reshape(n_out * D)
        elif T.is_torch(out):
            if len(shp) == 2:                        # y0 [B, D] -> [T*B, D] (time-major blocks)
                res = out.permute(1, 0, 2).reshape(n_out * B, D)
            else:                                    # y0 [..., L, D] -> [..., T*L, D]
                L = shp[-2]
                lead = shp[:-2]
                res = out.reshape(lead + (L, n_out, D)).transpose(-3, -2).reshape(lead + (n_out * L, D))
            res = res.contiguous()
        else:  # native buffers: the same two layouts on the host copy (numpy in -> numpy out anyway)
            h = T.to_host(out)
            if len(shp) == 2:
                h = np.ascontiguousarray(h.transpose(1, 0, 2).reshape(n_out * B, D))
            else:
                L = shp[-2]
                lead = shp[:-2]
                h = np.ascontiguousarray(np.swapaxes(h.reshape(lead + (L, n_out, D)), -3, -2).reshape(lead + (n_out * L, D)))
            res = h if isinstance(self.y0, (np.ndarray, list, tuple)) else T.from_host(h, out)
        return T.like_input(res, self.y0)


class Euler(FixedSolver):
    order = 1
    method = "euler"


class RK4(FixedSolver):
    order = 4
    method = "rk4"


class Midpoint(FixedSolver):  # fixed_solver/midpoint.py:4-18
    order = 2
    method = "midpoint"
