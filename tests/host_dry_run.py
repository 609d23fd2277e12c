"""Dry run of the Python HOST LAYER on a machine without a GPU -- TEST INFRASTRUCTURE ONLY.

`XDE_DRY_RUN=1 python -m pytest tests/<gpu test file> -m gpu` (tests/conftest.py installs this module when the variable
is set; tests/test_host_layer_dry_run.py does that in a subprocess as part of the CPU suite) executes GPU-marked tests
with two test doubles:

  * `libxde_b200.so` is replaced by `FakeLib`: every C-ABI entry point the selected tests reach is answered by the CPU
    oracle on the same raw pointers (host memory here);
  * a `TorchFunctionMode` sends every `device="cuda"` / `.cuda()` / `.to(cuda)` to the CPU and `torch.cuda.*` queries
    are stubbed.

What this checks is ONLY what runs above the C ABI: argument marshalling (pointer / size / struct order against
include/xde_b200.h), shapes and layouts, option routing, the autograd adapters, log and status decoding, and the test
code itself.  It says NOTHING about the CUDA kernels -- the numbers on both sides are the oracle's by construction --
and it is never active in a normal run: the product has no hook for it (the double is injected into the private module
global `paddlexde_b200._lib._lib` by the test session).  The parity tests proper are the same files run with `-m gpu`
on a B200."""
from __future__ import annotations

import ctypes as C

import numpy as np

PRE_NAME = {0: "id", 1: "square", 2: "cube"}
RK_NAME = {0: "dopri5", 1: "bosh3", 2: "fehlberg2", 3: "adaptive_heun", 4: "dopri8", 100: "dopri5"}
FIXED_NAME = {0: "euler", 1: "rk4", 2: "midpoint"}
INTERP_NAME = {0: "linear", 1: "cubic", 2: "bez"}
CTRL_NAME = {0: "trajectory", 1: "batch"}
ADJ_NORM_NAME = {0: "mixed", 1: "seminorm"}


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return int(p)


def _arr(p, shape, dtype=np.float32):
    """Writable numpy view of the raw buffer a C-ABI argument points to."""
    a = _addr(p)
    if not a:
        return None
    n = int(np.prod(shape)) if len(shape) else 1
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(a)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _obj(ref):
    return None if ref is None else ref._obj  # ctypes.byref(x)._obj is x


class FakeLib:
    """The entry points of include/xde_b200.h the dry-run tests reach, answered by the CPU oracle."""

    def __init__(self):
        from oracle import xde_oracle as xo

        xo.build()
        self.xo = xo
        self._err = ""
        self._launches = 0

    # ---- plumbing ----
    def xde_abi_version(self):
        return 3

    def xde_last_error(self):
        return self._err.encode()

    def xde_launch_count(self):
        return self._launches

    def xde_probe_ffma_f32(self, iters, sink, n_flops, stream):
        _obj(n_flops).value = int(iters) * 148 * 1024 * 16
        return 0

    xde_probe_ffma2_f32 = xde_probe_ffma_f32

    def __getattr__(self, name):
        def missing(*a, **k):
            raise NotImplementedError(f"host dry run: no double for {name}")

        if name.startswith("xde_"):
            return missing
        raise AttributeError(name)

    def _field(self, ref):
        f = _obj(ref)
        d, h = f.d, f.h
        return self.xo.MLP(_arr(f.w1, (d, h)).copy(), _arr(f.b1, (h,)).copy(), _arr(f.w2, (h, d)).copy(),
                           _arr(f.b2, (d,)).copy(), pre=PRE_NAME[f.pre])

    @staticmethod
    def _opts(ref):
        o = _obj(ref)
        return dict(rtol=o.rtol, atol=o.atol, min_step=o.min_step, max_step=o.max_step,
                    first_step=None if np.isnan(o.first_step) else o.first_step, safety=o.safety, ifactor=o.ifactor,
                    dfactor=o.dfactor, max_num_steps=o.max_num_steps)

    @staticmethod
    def _write_stats(p, stats, rc, B, controller):
        from paddlexde_b200._lib import StatsC

        a = _addr(p)
        if not a:
            return
        s = StatsC.from_address(a)
        mul = B if controller == "batch" else 1  # the batch kernels count every trajectory's attempt
        s.n_attempts = int(stats.n_attempts.sum()) * mul
        s.n_accepted = int(stats.n_accepted.sum()) * mul
        s.nfe = int(stats.nfe.sum()) * mul
        s.status = int(max(int(rc), int(stats.status.max())))

    @staticmethod
    def _write_log(log_ref, rows):
        """rows: {trajectory index: oracle log recarray}"""
        lg = _obj(log_ref)
        if lg is None or not rows:
            return
        nrow = max(rows) + 1
        cnt = _arr(lg.counts, (nrow,), np.int32)
        for b, r in rows.items():
            rec = _arr(C.c_void_p(lg.records + b * lg.cap * 16), (lg.cap, 4), np.float32)
            n = min(len(r), lg.cap)
            rec[:n, 0], rec[:n, 1], rec[:n, 2] = r.t0[:n], r.dt[:n], r.ratio[:n]
            rec[:n, 3] = np.asarray(r.accepted[:n], np.int32).view(np.float32)
            cnt[b] = len(r)

    # ---- adaptive forward ----
    def xde_adaptive_rk_mlp_grid_f32(self, method, field, y0, B, t_span, T, opts, controller, step_t, n_step, jump_t,
                                     n_jump, out, stats, log, stream):
        xo = self.xo
        self._launches += 1
        om, o = self._field(field), xo.make_opts(**self._opts(opts))
        D = om.d
        y0a, ta, outa = _arr(y0, (B, D)), _arr(t_span, (T,)), _arr(out, (T, B, D))
        ctrl = CTRL_NAME[controller]
        m = om.c()
        want = _obj(log) is not None
        rows = {}
        # forced points: the C ABI takes physical times in integration order, the oracle solver times (s = -t when the
        # span decreases, repair R5)
        sgn = np.float32(-1.0 if ta[1] < ta[0] else 1.0)
        st_a = np.ascontiguousarray(sgn * _arr(step_t, (n_step,))) if n_step else None
        jt_a = np.ascontiguousarray(sgn * _arr(jump_t, (n_jump,))) if n_jump else None
        ns = B if ctrl == "trajectory" else 1
        st = np.zeros(ns, xo.STATS_DTYPE)

        cap = (_obj(log).cap + 8) if want else 0  # the oracle keeps counting past the capacity, like the kernels
        rec = np.zeros(cap if (want and ctrl == "batch") else 0, xo.ATTEMPT_DTYPE)
        n = C.c_int64(0)
        rc = xo.lib().orc_adaptive_rk_mlp_grid(
            C.c_int32(xo.RK[RK_NAME[method]]), C.byref(m), xo._p(y0a), C.c_int64(B), xo._p(ta), C.c_int32(T),
            C.byref(o), C.c_int32(xo.CTRL[ctrl]), xo._p(st_a), C.c_int32(n_step), xo._p(jt_a), C.c_int32(n_jump),
            xo._p(outa), xo._p(st), xo._p(rec) if rec.size else None, C.c_int64(rec.size), C.c_int64(0), C.byref(n),
            C.c_int32(0))
        if want and ctrl == "batch":
            rows[0] = rec[:n.value].view(np.recarray)
        elif want:  # one controller per trajectory: every trajectory is its own solve, so its log is that of a B = 1 run
            for b in range(B):
                rec = np.zeros(cap, xo.ATTEMPT_DTYPE)
                n = C.c_int64(0)
                st1, out1 = np.zeros(1, xo.STATS_DTYPE), np.empty((T, 1, D), np.float32)
                yb = np.ascontiguousarray(y0a[b:b + 1])
                xo.lib().orc_adaptive_rk_mlp_grid(
                    C.c_int32(xo.RK[RK_NAME[method]]), C.byref(m), xo._p(yb), C.c_int64(1), xo._p(ta), C.c_int32(T),
                    C.byref(o), C.c_int32(xo.CTRL[ctrl]), xo._p(st_a), C.c_int32(n_step), xo._p(jt_a), C.c_int32(n_jump),
                    xo._p(out1), xo._p(st1), xo._p(rec), C.c_int64(rec.size), C.c_int64(0), C.byref(n), C.c_int32(0))
                rows[b] = rec[:n.value].view(np.recarray)
        self._write_stats(stats, st.view(np.recarray), rc, B, ctrl)
        self._write_log(log, rows)
        return 0

    def xde_adaptive_rk_mlp_f32(self, method, field, y0, B, t_span, T, opts, controller, out, stats, log, stream):
        return self.xde_adaptive_rk_mlp_grid_f32(method, field, y0, B, t_span, T, opts, controller, None, 0, None, 0, out,
                                                 stats, log, stream)

    def xde_dopri5_mlp_f32(self, field, y0, B, t_span, T, opts, controller, out, stats, log, stream):
        return self.xde_adaptive_rk_mlp_f32(0, field, y0, B, t_span, T, opts, controller, out, stats, log, stream)

    # ---- adjoint ----
    def xde_dopri5_mlp_adjoint_f32(self, field, t_span, T, y_ans, grad_y, B, opts, controller, adj_norm, out_g, out_a0,
                                   out_gt, stats, log, stream):
        xo = self.xo
        self._launches += 2  # the solve and the fp64 -> fp32 cast of the sums, as the library counts them
        om, kw = self._field(field), self._opts(opts)
        D = om.d
        ctrl, norm = CTRL_NAME[controller], ADJ_NORM_NAME[adj_norm]
        if ctrl == "batch" and _addr(out_gt):
            self._err = "grad_t_span is computed by the per-trajectory controller only"
            return -2
        if ctrl == "trajectory" and norm != "seminorm":
            self._err = "adjoint with one controller per trajectory supports the seminorm only"
            return -2
        ta, ya, ga = _arr(t_span, (T,)), _arr(y_ans, (T, B, D)), _arr(grad_y, (T, B, D))
        gt = np.zeros(T, np.float32) if _addr(out_gt) else None
        want = _obj(log) is not None
        rows = {}
        cap = (_obj(log).cap + 8) if want else 8
        g, a0, st, lg, rc = xo.dopri5_mlp_adjoint(om, ta, ya, ga, controller=ctrl, adj_norm=norm, log_cap=cap,
                                                  log_traj=0 if (want and ctrl == "batch") else None,
                                                  grad_t=None if gt is None else gt, **kw)
        if want and ctrl == "batch":
            rows[0] = lg
        elif want:  # per-trajectory controller: the log of trajectory b is the log of its own B = 1 solve
            for b in range(B):
                _, _, _, lgb, _ = xo.dopri5_mlp_adjoint(om, ta, ya[:, b:b + 1], ga[:, b:b + 1], controller=ctrl,
                                                        adj_norm=norm, log_traj=0, log_cap=cap,
                                                        grad_t=None if gt is None else np.zeros(T, np.float32), **kw)
                rows[b] = lgb
        _arr(out_g, (om.n_params,))[:] = g
        if _addr(out_a0):
            _arr(out_a0, (B, D))[:] = a0
        if gt is not None:
            _arr(out_gt, (T,))[:] = gt
        self._write_stats(stats, st, rc, B, ctrl)
        self._write_log(log, rows)
        return 0

    # ---- fixed grid ----
    def xde_rk_fixed_mlp_f32(self, method, field, y0, B, t_span, T, out_stride_t, out, stream):
        self._launches += 1
        om = self._field(field)
        D = om.d
        y = self.xo.fixed_mlp(FIXED_NAME[method], om, _arr(y0, (B, D)), _arr(t_span, (T,)))
        idx = list(range(0, T - 1, out_stride_t)) + [T - 1]
        _arr(out, (B, len(idx), D))[:] = y[:, idx]
        return 0

    def xde_fixed_interp_linear_f32(self, y_grid, grid, t_out, B, T, D, out, stream):
        self._launches += 1
        y, g, t, o = _arr(y_grid, (B, T, D)), _arr(grid, (T,)), _arr(t_out, (T,)), _arr(out, (B, T, D))
        o[:, 0] = y[:, 0]
        for i in range(1, T):  # linear_interp, interpolation/functional/interp_fn.py:4-10
            if t[i] == g[i - 1]:
                o[:, i] = y[:, i - 1]
            elif t[i] == g[i]:
                o[:, i] = y[:, i]
            else:
                slope = np.float32(np.float32(t[i] - g[i - 1]) / np.float32(g[i] - g[i - 1]))
                o[:, i] = y[:, i - 1] + slope * (y[:, i] - y[:, i - 1])
        return 0

    # ---- SDE with a supplied table; the double of a tensor-core entry is the FP32 oracle (status word: in range) ----
    def xde_rk_fixed_mlp_tc_f32(self, method, field, y0, B, t_span, T, out_stride_t, out, status, stream):
        if _addr(status):
            _arr(status, (1,), np.int32)[0] = 0
        return self.xde_rk_fixed_mlp_f32(method, field, y0, B, t_span, T, out_stride_t, out, stream)

    def xde_sde_mlp_tc_f32(self, scheme, drift, diffusion, y0, B, t_span, T, dW, out_stride_t, out, status, stream):
        if scheme != 0:
            self._err = "the tensor-core SDE kernel integrates Euler-Maruyama only"
            return -2
        if _addr(status):
            _arr(status, (1,), np.int32)[0] = 0
        return self.xde_sde_mlp_f32(scheme, drift, diffusion, y0, B, t_span, T, dW, out_stride_t, out, stream)

    def xde_sde_mlp_f32(self, scheme, drift, diffusion, y0, B, t_span, T, dW, out_stride_t, out, stream):
        self._launches += 1
        f, g = self._field(drift), self._field(diffusion)
        D = f.d
        y = self.xo.sde_mlp({0: "em", 1: "milstein"}[scheme], f, g, _arr(y0, (B, D)), _arr(t_span, (T,)),
                            _arr(dW, (T - 1, B, D)))
        idx = list(range(0, T - 1, out_stride_t)) + [T - 1]
        _arr(out, (B, len(idx), D))[:] = y[:, idx]
        return 0

    def xde_brownian_increments_f32(self, seed, traj_offset, t_span, T, B, D, dW, stream):
        from oracle import philox_np

        self._launches += 1
        _arr(dW, (T - 1, B, D))[:] = philox_np.brownian_increments(int(seed), _arr(t_span, (T,)), B, D, int(traj_offset))
        return 0

    def xde_sde_mlp_philox_f32(self, scheme, math, drift, diffusion, y0, B, t_span, T, seed, traj_offset, out_stride_t,
                               out, status, stream):
        from oracle import philox_np

        if math == 1 and scheme != 0:
            self._err = "the tensor-core SDE kernel integrates Euler-Maruyama only"
            return -2
        if _addr(status):
            _arr(status, (1,), np.int32)[0] = 0
        dW = philox_np.brownian_increments(int(seed), _arr(t_span, (T,)), B, self._field(drift).d, int(traj_offset))
        return self.xde_sde_mlp_f32(scheme, drift, diffusion, y0, B, t_span, T, C.c_void_p(dW.ctypes.data), out_stride_t,
                                    out, stream)

    def xde_sde_mlp_adjoint_f32(self, drift, diffusion, t_span, T, y_all, grad_y, B, dW, seed, traj_offset, out_gf, out_gg,
                                out_a0, stream):
        from oracle import philox_np

        self._launches += 1
        f, g = self._field(drift), self._field(diffusion)
        D = f.d
        ta = _arr(t_span, (T,))
        table = _arr(dW, (T - 1, B, D)) if _addr(dW) else philox_np.brownian_increments(int(seed), ta, B, D, int(traj_offset))
        gf, gg, a0 = self.xo.sde_mlp_adjoint(f, g, ta, _arr(y_all, (B, T, D)), _arr(grad_y, (B, T, D)), table)
        _arr(out_gf, (f.n_params,))[:] = gf
        _arr(out_gg, (g.n_params,))[:] = gg
        if _addr(out_a0):
            _arr(out_a0, (B, D))[:] = a0
        return 0

    # ---- delay path ----
    def xde_history_gather_f32(self, kind, his, R, Th, D, span, lags, L, out_val, out_der, stream):
        self._launches += 1
        if Th < 2 or (kind == 2 and Th < 4):
            self._err = "history gather: too few samples"
            return -1
        v, d = self.xo.history_gather(INTERP_NAME[kind], _arr(his, (R, Th, D)), _arr(span, (Th,)), _arr(lags, (L,)))
        _arr(out_val, (R, L, D))[:] = v
        _arr(out_der, (R, L, D))[:] = d
        return 0

    def xde_history_gather_bwd_f32(self, grad_y, deriv, R, L, D, out, stream):
        self._launches += 1
        _arr(out, (L,))[:] = self.xo.history_gather_bwd(_arr(grad_y, (R, L, D)), _arr(deriv, (R, L, D)))
        return 0

    def xde_dde_fuse_f32(self, dy, dt, y0, n, y1, stream):
        self._launches += 1
        _arr(y1, (n,))[:] = self.xo.dde_fuse(_arr(dy, (n,)), float(dt), _arr(y0, (n,)))
        return 0

    def xde_dde_fuse_bwd_f32(self, grad_y1, dt, n, grad_dy, grad_y0, stream):
        # y1 = (dy - 0.001 (dy dt + y0)) dt + y0:  d/d dy = dt (1 - 0.001 dt),  d/d y0 = 1 - 0.001 dt
        self._launches += 1
        g, dt = _arr(grad_y1, (n,)), np.float32(dt)
        c = np.float32(1.0) - np.float32(0.001) * dt
        if _addr(grad_dy):
            _arr(grad_dy, (n,))[:] = g * (dt * c)
        if _addr(grad_y0):
            _arr(grad_y0, (n,))[:] = g * c
        return 0


def install():
    """Route CUDA placement to the CPU and the C ABI to the oracle, for this process."""
    import torch
    from torch.overrides import TorchFunctionMode

    def is_cuda_dev(d):
        return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    class CpuAsCuda(TorchFunctionMode):
        def __torch_function__(self, func, types, args=(), kwargs=None):
            kwargs = dict(kwargs or {})
            if is_cuda_dev(kwargs.get("device")):
                kwargs["device"] = "cpu"
            if func is torch.Tensor.cuda:
                return args[0]
            if func is torch.Tensor.to:
                args = tuple("cpu" if is_cuda_dev(a) else a for a in args)
                kwargs.pop("non_blocking", None)
            if func is torch.Tensor.pin_memory:
                return args[0]
            return func(*args, **kwargs)

    mode = CpuAsCuda()
    mode.__enter__()

    import contextlib
    import time
    import types

    class _Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def wait_stream(self, s):
            pass

        def wait_event(self, e):
            pass

        def synchronize(self):
            pass

    class _Event:  # host wall clock: the dry run times nothing meaningful, it only has to produce numbers
        def __init__(self, *a, **k):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return max((other.t - self.t) * 1e3, 1e-3)

        def synchronize(self):
            pass

        def query(self):
            return True

        def wait(self, *a):
            pass

    @contextlib.contextmanager
    def _stream_ctx(s):
        yield

    def _no_graph(*a, **k):
        raise RuntimeError("host dry run: no CUDA graphs")

    torch.cuda.Stream = _Stream
    torch.cuda.Event = _Event
    torch.cuda.stream = _stream_ctx
    torch.cuda.CUDAGraph = _no_graph
    torch.cuda.get_device_properties = lambda *a, **k: types.SimpleNamespace(
        uuid="host-dry-run", multi_processor_count=148, name="host dry run", total_memory=180 << 30)
    torch.cuda.is_current_stream_capturing = lambda: False  # asked by torch.optim before every step
    torch.cuda.is_available = lambda: True
    torch.cuda.current_device = lambda: 0
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None

    from paddlexde_b200 import _lib, _tensor, distributed

    init_from_env = distributed.init_from_env
    distributed.init_from_env = lambda backend=None: init_from_env("gloo")  # no NCCL without GPUs
    _tensor.device = lambda: torch.device("cpu")  # autograd's backward thread runs outside the function mode
    _lib._lib = FakeLib()
    return mode
