"""Problem wrappers with the reference's protocol (paddlexde/xde/base_xde.py:8-107).

In the reference the solver calls back ``xde.move`` / ``xde.fuse`` once per stage.  Here the whole
stepping loop runs inside one CUDA kernel, so a wrapper *describes* the problem (which fused field,
which update rule) and the solver dispatches on that description.  ``move``/``fuse`` remain as the
protocol names; calling them from Python is not part of the hot path."""
from __future__ import annotations


class BaseXDE:
    def __init__(self, name, var_nums, y0, t_span):
        self.name = name
        self.var_nums = var_nums
        self.t_span = t_span
        self.pred_len = getattr(t_span, "shape", (len(t_span),))

    def method(self):
        return self.name

    def init_y0(self, y0):
        self.y0 = y0

    def on_integrate_step_end(self, y0=None, y1=None, t0=None, t1=None):
        pass

    def format(self, sol):
        """Repair R1: `xde.format` is called by odeint (functional/odeint.py:33) but only
        exists commented out (base_xde.py:89-100); identity on what `integrate` returns."""
        return sol
