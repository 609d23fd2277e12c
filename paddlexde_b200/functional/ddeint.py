from __future__ import annotations

import torch

from .. import _tensor as T
from ..utils.ode_utils import _rms_norm
from ..xde import BaseDDE


def ddeint(func, y0, t_span, lags, his, his_span, solver, his_processed=False, rtol=1e-7, atol=1e-9,
           options: object = {"norm": _rms_norm}, fixed_solver_interp="linear"):
    """Same signature as paddlexde/functional/ddeint.py:9-47; returns (solution, xde.y_lags).

    The history resampling (HistoryIndex) and the damped-Euler fuse run as CUDA kernels; `func(y_lags,
    y0)` -- in D3STN a full transformer, out of scope here -- is the caller's callable and is invoked
    once per step exactly as FixedSolver.integrate does (solver/base_fixed_solver.py:125-141)."""
    from ..solver import Euler

    xde = BaseDDE(func, y0=y0, t_span=t_span, lags=lags, his=his, his_span=his_span, his_processed=his_processed)
    if solver is not Euler:
        raise NotImplementedError("ddeint is fused for solver=Euler (the D3STN configuration, train_dde.py:418-433)")
    t = torch.as_tensor(t_span, dtype=torch.float32).reshape(-1).cpu()
    y = xde.y0
    sol = [y]
    for i in range(1, t.numel()):
        dt = float(t[i] - t[i - 1])
        dy = xde.move(t[i - 1], dt, y)            # Euler.step fixed_solver/euler.py:7-11
        y = xde.fuse(dy, dt, y)
        sol.append(y)
    solution = torch.cat([T.to_dev(s) for s in sol], dim=-2)
    return solution, xde.y_lags
