"""The fused vector-field family.

The reference accepts any ``func(t, y)`` (xde/base_ode.py:60-62).  A fused kernel cannot run an
arbitrary Python callable, so the supported family is declared explicitly:

    f(t, y) = tanh(pre(y) @ W1 + b1) @ W2 + b2          (example/ode_demo.py:17-33)

with Paddle ``nn.Linear`` weight layout ``[in, out]`` and ``pre`` one of ``id | square | cube``
(``y**3`` in ode_demo, ``y**2`` in sde_demo's diffusion).  Anything else is a hard
``UnsupportedFieldError`` -- there is no fallback path."""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import _tensor as T
from ._lib import PRE, MlpFieldC, UnsupportedFieldError


class MLPField:
    def __init__(self, w1, b1, w2, b2, pre: str = "cube", act: str = "tanh", weight_layout: str = "in_out"):
        """weight_layout: "in_out" = Paddle nn.Linear ([in, out]: w1 [D, H], w2 [H, D]); "out_in" = torch nn.Linear
        ([out, in]: w1 [H, D], w2 [D, H]).  The caller's tensors are kept as they are (they stay the leaves that
        autograd / the optimizer see); the kernels read [in, out] device copies made here and by refresh()."""
        if weight_layout not in ("in_out", "out_in"):
            raise ValueError("weight_layout must be 'in_out' (Paddle) or 'out_in' (torch)")
        self.weight_layout = weight_layout
        if act != "tanh":
            raise UnsupportedFieldError(f"activation {act!r} has no fused kernel (tanh only)")
        if pre not in PRE:
            raise UnsupportedFieldError(f"pre-activation {pre!r} has no fused kernel (id|square|cube)")
        self.pre = pre
        self._src = (w1, b1, w2, b2)  # kept so autograd can route gradients to the caller's tensors
        self._load()
        if self.w1.dim() != 2:
            raise ValueError("w1 must be [D, H] (Paddle nn.Linear layout [in, out])")
        self.d, self.h = self.w1.shape
        if tuple(self.w2.shape) != (self.h, self.d) or tuple(self.b1.shape) != (self.h,) \
                or tuple(self.b2.shape) != (self.d,):
            raise ValueError("field shapes must be w1[D,H], b1[H], w2[H,D], b2[D]")

    # -- reference-facing surface ---------------------------------------------------------------
    def parameters(self) -> Sequence:
        """Like nn.Layer.parameters() (functional/odeint_adjoint.py:276)."""
        return list(self._src)

    def _load(self):
        def dev(a, transpose):
            if T.is_torch(a):
                t = T.to_dev(a)
                return t.t().contiguous() if transpose else t
            h = np.asarray(T.to_host(a), dtype=np.float32)
            return T.to_dev(np.ascontiguousarray(h.T) if transpose else h)

        tr = self.weight_layout == "out_in"
        w1, b1, w2, b2 = self._src
        self.w1, self.b1, self.w2, self.b2 = dev(w1, tr), dev(b1, False), dev(w2, tr), dev(b2, False)
        self._versions = self._src_versions()

    def _src_versions(self):
        # torch bumps `_version` on every in-place update (optimizer.step()): lets c_struct() notice stale device copies
        return tuple(getattr(a, "_version", None) for a in self._src)

    def refresh(self):
        """Re-read the caller's parameter tensors.  torch tensors are re-read automatically when an in-place update
        (optimizer.step()) has bumped their version counter; call this after changing numpy / DeviceArray sources
        (those are copied at construction) or after rebinding `.data`."""
        self._load()
        return self

    def grads_like_params(self, flat):
        """The flat (gW1, gb1, gW2, gb2) of the adjoint entries, shaped like `parameters()`."""
        gw1, gb1, gw2, gb2 = self.split_flat(flat)
        if self.weight_layout == "out_in":
            gw1, gw2 = gw1.t(), gw2.t()
        return [gw1, gb1, gw2, gb2]

    @property
    def n_params(self) -> int:
        return 2 * self.d * self.h + self.h + self.d

    def split_flat(self, flat):
        d, h = self.d, self.h
        out, o = [], 0
        for shp in ((d, h), (h,), (h, d), (d,)):
            n = 1
            for v in shp:
                n *= v
            out.append(flat[o:o + n].reshape(shp))
            o += n
        return out

    def c_struct(self) -> MlpFieldC:
        if self._src_versions() != self._versions:  # the caller's tensors were updated in place since the last load
            self._load()
        return MlpFieldC(self.d, self.h, PRE[self.pre], 0, self.w1.data_ptr(), self.b1.data_ptr(),
                         self.w2.data_ptr(), self.b2.data_ptr())

    # -- extraction from a framework module -------------------------------------------------------
    @classmethod
    def from_sequential(cls, net, pre: str = "cube", weight_layout: str = "auto") -> "MLPField":
        """Recognise ``Sequential(Linear, Tanh, Linear)`` of Paddle or torch; hard-fail otherwise."""
        layers = list(net.children()) if hasattr(net, "children") else list(net)
        names = [type(l).__name__ for l in layers]
        if names != ["Linear", "Tanh", "Linear"]:
            raise UnsupportedFieldError(f"only Sequential(Linear, Tanh, Linear) is fused; got {names}")
        l1, _, l2 = layers
        w1, w2 = l1.weight, l2.weight
        if weight_layout == "auto":
            weight_layout = "out_in" if T.is_torch(w1) else "in_out"
        # the module's own Parameters stay the leaves: gradients are routed (and transposed back) to them, and
        # refresh() re-reads them after an optimizer step
        return cls(w1, l1.bias, w2, l2.bias, pre=pre, weight_layout=weight_layout)


def as_field(func) -> MLPField:
    """What `odeint(func, ...)` accepts: an MLPField, or an object carrying one as `.xde_field`."""
    if isinstance(func, MLPField):
        return func
    f = getattr(func, "xde_field", None)
    if isinstance(f, MLPField):
        return f
    raise UnsupportedFieldError(
        "paddlexde_b200 integrates the fused field family only: pass an MLPField (or an object with an "
        "`.xde_field` MLPField attribute, e.g. MLPField.from_sequential(func.net, pre='cube')). "
        "Arbitrary Python callables cannot run inside the CUDA stepper and there is no fallback.")
