from __future__ import annotations

import numpy as np

from .. import _tensor as T
from .._tensor import torch
from ..utils.ode_utils import _rms_norm
from ..xde import BaseDDE


_one_third = 1 / 3
_two_thirds = 2 / 3


def _step(method, xde, t0, t1, y0):
    """FixedSolver.step of the reference driven through xde.move / xde.fuse
    (fixed_solver/euler.py:7-11, midpoint.py:7-18, rk4.py:7-10 -> rk4_alt_step_func base_fixed_solver.py:166-197).
    The time arithmetic is the reference's: fp32 tensors (`dt * _one_third` is an fp32 product of two fp32 values)."""
    t0, t1 = np.float32(t0), np.float32(t1)
    dt = np.float32(t1 - t0)
    if method == "euler":
        dy = xde.move(t0, dt, y0)
        return xde.fuse(dy, dt, y0), dy
    if method == "midpoint":
        half_dt = np.float32(0.5) * dt
        y_half = xde.fuse(xde.move(t0, half_dt, y0), half_dt, y0)
        dy = xde.move(t0 + half_dt, dt, y_half)
        return xde.fuse(dy, dt, y0), dy
    if method == "rk4":
        k1 = xde.move(t0, dt, y0)
        dt3 = dt * np.float32(_one_third)
        k2 = xde.move(t0 + dt3, dt3, xde.fuse(k1, dt3, y0))
        k3 = xde.move(t0 + dt * np.float32(_two_thirds), dt3, xde.fuse(k1 - k2 * _one_third, dt, y0))
        k4 = xde.move(t1, t0 + dt3, xde.fuse(k1 - k2 + k3, dt, y0))
        y1 = (xde.fuse(k1, dt, y0) + 3 * xde.fuse(k2, dt, y0) + 3 * xde.fuse(k3, dt, y0) + xde.fuse(k4, dt, y0)) * 0.125
        return y1, k1
    raise NotImplementedError(f"ddeint: no fixed-step scheme {method!r}")


def ddeint(func, y0, t_span, lags, his, his_span, solver, his_processed=False, rtol=1e-7, atol=1e-9,
           options: object = {"norm": _rms_norm}, fixed_solver_interp="linear"):
    """Same signature as paddlexde/functional/ddeint.py:9-47; returns (solution, xde.y_lags).

    The history resampling (HistoryIndex) and the damped fuse run as CUDA kernels (both differentiable: the
    solution stays on the autograd graph, as the D3STN trainer needs, example/D3STN/train_dde.py:424-454);
    `func(y_lags, y0)` -- in D3STN a full transformer, out of scope here -- is the caller's callable and is
    invoked once per stage exactly as FixedSolver.integrate does (solver/base_fixed_solver.py:125-141).
    solver: Euler (the D3STN configuration, train_dde.py:418-433), Midpoint or RK4; grid == t_span, where both
    output interpolants return the end of the step (interp_fn.py:4-20)."""
    from ..solver import FixedSolver
    from ..xde.base_dde import _as_graph_tensor

    if torch is None:
        raise ImportError("ddeint calls back into the caller's `func(y_lags, y)` on framework tensors: it needs PyTorch "
                          "(the history gather and the fuse kernels are available without it: xde.base_dde)")

    xde = BaseDDE(func, y0=y0, t_span=t_span, lags=lags, his=his, his_span=his_span, his_processed=his_processed)
    if not (isinstance(solver, type) and issubclass(solver, FixedSolver)):
        raise NotImplementedError("ddeint integrates with the fixed-step solvers (Euler | Midpoint | RK4); the "
                                  "reference's D3STN uses Euler (train_dde.py:418-433)")
    if fixed_solver_interp not in ("linear", "cubic", "", None):
        raise ValueError(f"interp must be 'linear' or 'cubic', got {fixed_solver_interp!r}")
    t = torch.as_tensor(t_span, dtype=torch.float32).reshape(-1).cpu()
    y = _as_graph_tensor(xde.y0)
    sol = [y]
    for i in range(1, t.numel()):
        y, _ = _step(solver.method, xde, float(t[i - 1]), float(t[i]), y)
        xde.on_integrate_step_end(y0=sol[-1], y1=y, t0=t[i - 1], t1=t[i])
        sol.append(y)
    solution = torch.cat(sol, dim=-2)  # concat(axis=-2), base_fixed_solver.py:143
    return solution, xde.y_lags
