from __future__ import annotations

from ..utils.ode_utils import _rms_norm
from ..xde import BaseODE


def odeint(func, y0, t_span, solver, *, rtol=1e-7, atol=1e-9, options: object = {"norm": _rms_norm}):
    """Same signature and flow as paddlexde/functional/odeint.py:9-35."""
    xde = BaseODE(func, y0=y0, t_span=t_span)
    s = solver(xde=xde, y0=xde.y0, rtol=rtol, atol=atol, **options)
    solution = s.integrate(t_span)
    solution = xde.format(solution)
    odeint.last_solver = s
    return solution
