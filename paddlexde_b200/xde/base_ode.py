from __future__ import annotations

from ..field import as_field
from .base_xde import BaseXDE


class BaseODE(BaseXDE):
    """paddlexde/xde/base_ode.py:9-62: move = func(t, y); fuse = dy*dt + y0."""
    kind = "ode"

    def __init__(self, func, y0, t_span):
        super().__init__(name="ODE", var_nums=1, y0=y0, t_span=t_span)
        self.func = func
        self.field = as_field(func)  # hard error for an unsupported field
        self.init_y0(y0)
