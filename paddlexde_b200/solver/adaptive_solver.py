"""The embedded Runge-Kutta solvers (Dopri5, Bosh3, Fehlberg2, AdaptiveHeun, Dopri8) with the
reference's solver-class protocol:

    s = Dopri5(xde=xde, y0=xde.y0, rtol=rtol, atol=atol, **options); s.integrate(t_span)

(paddlexde/functional/odeint.py:30-31; constructor keywords solver/base_adaptive_solver_rk.py:32-49).
`integrate` is ONE kernel launch (csrc/xde_dopri5_fwd.cu for Dopri5, the table-driven
csrc/xde_adaptive_rk.cu for the other tableaux); the controller lives on the device."""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from dataclasses import dataclass

import numpy as np

from .. import _tensor as T
from .._lib import (CTRL, RK, AttemptLogC, CtrlOptsC, StatsC, UnsupportedFieldError, check, lib,
                    raise_for_status)


class ControllerDefaultWarning(UserWarning):
    """Raised once per process when the controller granularity was not chosen by the caller (see default_controller)."""


_warned_default = False


def default_controller(batch_rows: int | None = None) -> str:
    """Controller granularity when the caller does not pass `controller=`.

    The reference applies `_rms_norm` to the whole [B, D] state: ONE dt and one accept/reject for the batch, and its
    adjoint default is the mixed norm over (y, a, g_theta) (utils/ode_utils.py:8-9, functional/odeint_adjoint.py:
    284-291) -- that is controller="batch".  This package defaults to one controller per trajectory
    ("trajectory": each trajectory is stepped as the reference would step it alone, B = 1; the adjoint then uses the
    seminorm), which is what BASELINE.json's north star asks for and what the throughput kernels implement.  For
    B > 1 the two give different step sequences (both within tolerance), so the choice is announced once per process.
    Set PADDLEXDE_B200_CONTROLLER=batch|trajectory to choose the default globally (no warning then)."""
    global _warned_default
    env = os.environ.get("PADDLEXDE_B200_CONTROLLER")
    if env:
        if env not in CTRL:
            raise ValueError("PADDLEXDE_B200_CONTROLLER must be 'trajectory' or 'batch'")
        return env
    if not _warned_default and (batch_rows is None or batch_rows > 1):
        _warned_default = True
        warnings.warn("paddlexde_b200 steps every trajectory with its own error controller (controller='trajectory', "
                      "adjoint seminorm); the reference uses one global RMS norm / dt for the whole batch and the mixed "
                      "adjoint norm.  Pass options={'controller': 'batch'} (or set PADDLEXDE_B200_CONTROLLER=batch) for "
                      "the reference's step sequence; pass controller='trajectory' to silence this notice.",
                      ControllerDefaultWarning, stacklevel=3)
    return "trajectory"


@dataclass
class SolveStats:
    n_attempts: int = 0
    n_accepted: int = 0
    nfe: int = 0
    status: int = 0


def make_ctrl_opts(rtol, atol, min_step=0.0, max_step=float("inf"), first_step=None, safety=0.9,
                   ifactor=10.0, dfactor=0.2, max_num_steps=2 ** 31 - 1) -> CtrlOptsC:
    return CtrlOptsC(float(rtol), float(atol), float(min_step), float(max_step),
                     float("nan") if first_step is None else float(first_step), float(safety),
                     float(ifactor), float(dfactor), int(min(max_num_steps, 2 ** 31 - 1)), 0)


def check_norm(norm):
    """options["norm"] is a selector here (utils/ode_utils.py): only the RMS norm is fused."""
    if norm is None or getattr(norm, "xde_norm", None) == "rms":
        return
    raise UnsupportedFieldError("only the reference's default `_rms_norm` error norm is fused on the device; "
                                "a Python norm callable cannot run inside the CUDA controller")


def host_tspan(t_span) -> np.ndarray:
    """fp32 host copy of t_span (solver/base_adaptive_solver.py:27 casts to the fp32 time dtype)."""
    t = np.asarray(T.to_host(t_span), dtype=np.float32)
    t = np.ascontiguousarray(t.reshape(-1))
    d = np.diff(t)
    if t.size < 2 or not (np.all(d > 0) or np.all(d < 0)):
        raise ValueError("t_span must be strictly increasing or strictly decreasing with at least 2 points")
    return t


_STATS_WORDS = C.sizeof(StatsC) // 8


def _decode_stats(words) -> SolveStats:
    s = StatsC.from_buffer_copy(words.tobytes())
    return SolveStats(int(s.n_attempts), int(s.n_accepted), int(s.nfe), int(s.status))


class StatsBuffer:
    """Device-resident xde_stats_t read back lazily (one tiny D2H copy when asked).  The entry points zero it."""

    def __init__(self, like, buf=None):
        """like: any device buffer of the call (decides the provider: torch / native)."""
        self.buf = T.empty((_STATS_WORDS,), like, "i64") if buf is None else buf

    def read(self) -> SolveStats:
        return _decode_stats(T.to_host(self.buf))


class StatsPair:
    """The forward solve's and the adjoint solve's xde_stats_t side by side, so that odeint_adjoint's backward reads
    both status words with ONE device-to-host copy."""

    def __init__(self, like):
        self.both = T.zeros((2 * _STATS_WORDS,), like, "i64")
        self.fwd = StatsBuffer(like, self.both[:_STATS_WORDS])
        self.adj = StatsBuffer(like, self.both[_STATS_WORDS:])

    def read(self):
        w = T.to_host(self.both)
        return _decode_stats(w[:_STATS_WORDS]), _decode_stats(w[_STATS_WORDS:])


_tspan_cache: dict = {}


def device_tspan(t_host: np.ndarray, like):
    """fp32 device copy of t_span (same provider and device as `like`); the handful of grids a training loop uses are
    cached (forward and backward ask for the same one every step, and a pageable H2D copy costs more than the rest of
    the launch path)."""
    key = (t_host.tobytes(), str(getattr(like, "device", None)), T.is_torch(like))
    t = _tspan_cache.get(key)
    if t is None:
        if len(_tspan_cache) >= 64:
            _tspan_cache.clear()
        t = T.from_host(t_host, like)
        _tspan_cache[key] = t
    return t


class AttemptLog:
    """Optional per-trajectory attempt log (tests / diagnostics): records [B, cap]."""
    dtype = np.dtype([("t0", np.float32), ("dt", np.float32), ("ratio", np.float32), ("accepted", np.int32)])

    def __init__(self, B, cap, like):
        self.cap = cap
        self.records = T.zeros((B, cap, 4), like, "f32")
        self.counts = T.zeros((B,), like, "i32")

    def c_struct(self):
        return AttemptLogC(self.records.data_ptr(), self.counts.data_ptr(), self.cap, 0)

    def read(self):
        rec = T.to_host(self.records).view(self.dtype).reshape(self.records.shape[0], self.cap)
        return rec.view(np.recarray), T.to_host(self.counts)


class AdaptiveRKSolver:
    """solver/base_adaptive_solver_rk.py:26-49; subclasses name the tableau (`method`) like the reference's
    adaptive_solver/*.py name theirs."""
    order = 5
    method = "dopri5"

    def __init__(self, xde, y0, rtol, atol, min_step=0, max_step=float("inf"), first_step=None, step_t=None,
                 jump_t=None, safety=0.9, ifactor=10.0, dfactor=0.2, max_num_steps=2 ** 31 - 1, dtype=None,
                 norm=None, controller=None, log_attempts=0, check_status=True, stats_buffer=None, **unused_kwargs):
        self.step_t, self.jump_t = step_t, jump_t
        if getattr(xde, "kind", None) != "ode":
            raise UnsupportedFieldError(f"{type(self).__name__} integrates BaseODE problems")
        check_norm(norm)
        if controller is None:
            shp = getattr(y0, "shape", None)
            rows = None
            if shp is not None and len(shp) >= 1:
                rows = 1
                for v in shp[:-1]:
                    rows *= int(v)
            controller = default_controller(rows)
        if controller not in CTRL:
            raise ValueError("controller must be 'trajectory' or 'batch'")
        self.xde, self.y0 = xde, y0
        self.rtol, self.atol = rtol, atol
        self.opts = make_ctrl_opts(rtol, atol, min_step, max_step, first_step, safety, ifactor, dfactor,
                                   max_num_steps)
        self.controller = controller
        self.log_attempts = int(log_attempts)
        self.check_status = check_status
        self._stats_buf = stats_buffer  # optional caller-owned StatsBuffer (odeint_adjoint shares one with its backward)
        self.stats = None
        self.attempt_log = None

    def integrate(self, t_span):
        """-> [T, *y0.shape] (solver/base_adaptive_solver.py:25-31)."""
        field = self.xde.field
        y0 = T.to_dev(self.y0)
        if y0.shape[-1] != field.d:
            raise ValueError(f"y0 last dim {y0.shape[-1]} != field state dim {field.d}")
        B = y0.numel() // field.d
        t_host = host_tspan(t_span)
        t_dev = device_tspan(t_host, y0)
        out = T.empty((t_host.size,) + tuple(y0.shape), y0)
        if self._stats_buf is None:
            self._stats_buf = StatsBuffer(y0)
        log_c = None
        if self.log_attempts > 0:
            self.attempt_log = AttemptLog(B, self.log_attempts, y0)
            log_c = C.byref(self.attempt_log.c_struct())
        fs = field.c_struct()
        step_t, jump_t = (self._sort_tvals(v, t_host, y0) for v in (self.step_t, self.jump_t))
        check(lib().xde_adaptive_rk_mlp_grid_f32(
            RK[self.method], C.byref(fs), T.ptr(y0), B, T.ptr(t_dev), t_host.size, C.byref(self.opts),
            CTRL[self.controller], T.ptr(step_t) if step_t is not None else None, 0 if step_t is None else step_t.numel(),
            T.ptr(jump_t) if jump_t is not None else None, 0 if jump_t is None else jump_t.numel(), T.ptr(out),
            T.ptr(self._stats_buf.buf), log_c, T.stream(y0)))
        if self.check_status:  # the reference asserts synchronously; opt out to stay asynchronous
            self.stats = self._stats_buf.read()
            raise_for_status(self.stats.status)
        return T.like_input(out, self.y0)

    @staticmethod
    def _sort_tvals(tvals, t_host, like):
        """sort_tvals (utils/ode_utils.py:22-25): drop the points before t_span[0], sort in integration
        order (ascending for an increasing t_span, descending for a decreasing one: repair R5)."""
        if tvals is None:
            return None
        v = np.asarray(T.to_host(tvals)).astype(np.float32).reshape(-1)
        if t_host[1] < t_host[0]:
            v = np.sort(v[v <= t_host[0]])[::-1]
        else:
            v = np.sort(v[v >= t_host[0]])
        if v.size == 0:
            return None
        return T.from_host(v.copy(), like)  # copy(): a reversed length-1 view keeps its negative stride

    def read_stats(self) -> SolveStats:
        self.stats = self._stats_buf.read()
        return self.stats


class Dopri5(AdaptiveRKSolver):          # adaptive_solver/dopri5.py:58-61
    order, method = 5, "dopri5"


class Bosh3(AdaptiveRKSolver):           # adaptive_solver/bosh3.py:24-27
    order, method = 3, "bosh3"


class Fehlberg2(AdaptiveRKSolver):       # adaptive_solver/fehlberg2.py:19-22
    order, method = 2, "fehlberg2"


class AdaptiveHeun(AdaptiveRKSolver):    # adaptive_solver/adaptive_heun.py:24-27
    order, method = 2, "adaptive_heun"


class Dopri8(AdaptiveRKSolver):          # adaptive_solver/dopri8.py:249-252
    order, method = 8, "dopri8"


class _Dopri5Table(AdaptiveRKSolver):
    """Diagnostics: the table-driven kernel with the Dormand-Prince tableau (== Dopri5 bit for bit)."""
    order, method = 5, "dopri5_table"
