from __future__ import annotations

from .. import _tensor as T
from .._lib import INTERP, check, lib
from .._tensor import torch  # None when PyTorch is not installed: the gathers below still work
from .base_xde import BaseXDE


def history_gather(lags, his, his_span, interp_method="cubic"):
    """interp.evaluate(lags), interp.derivative(lags) in one gather kernel
    (interpolation/interpolate_base.py:49-114). his [..., Th, D] -> two [..., L, D] tensors."""
    if interp_method not in ("linear", "cubic", "bez"):
        raise NotImplementedError(interp_method)  # xde/base_dde.py:104-111
    his_d = T.to_dev(his, like=lags if T.is_torch(lags) else None)
    span_d, lags_d = T.to_dev(his_span, like=his_d), T.to_dev(lags, like=his_d).reshape(-1)
    lead, Th, D = his_d.shape[:-2], his_d.shape[-2], his_d.shape[-1]
    if span_d.numel() != Th:
        raise ValueError("his_span must have his.shape[-2] entries")
    R = 1
    for s in lead:
        R *= s
    L = lags_d.numel()
    val = T.empty(tuple(lead) + (L, D), his_d)
    der = T.empty(tuple(lead) + (L, D), his_d)
    check(lib().xde_history_gather_f32(INTERP[interp_method], T.ptr(his_d), R, Th, D, T.ptr(span_d),
                                       T.ptr(lags_d), L, T.ptr(val), T.ptr(der), T.stream(his_d)))
    return val, der


def history_gather_bwd(grad_y, deriv):
    """sum(grad_y * deriv, axis=[0,1,3]) generalised to all leading dims (xde/base_dde.py:121-127)."""
    g = T.to_dev(grad_y, like=deriv if T.is_torch(deriv) else None)
    d = T.to_dev(deriv, like=g)
    L, D = g.shape[-2], g.shape[-1]
    R = g.numel() // (L * D)
    out = T.empty((L,), g)
    check(lib().xde_history_gather_bwd_f32(T.ptr(g), T.ptr(d), R, L, D, T.ptr(out), T.stream(g)))
    return out


class HistoryIndex(torch.autograd.Function if torch is not None else object):
    """xde/base_dde.py:82-127: differentiable (wrt the real-valued lags) resampling of a fixed history."""

    @staticmethod
    def forward(ctx, lags, his, his_span, interp_method="cubic"):
        y_lags, deriv = history_gather(lags, his, his_span, interp_method)
        ctx.save_for_backward(deriv)
        return y_lags

    @classmethod
    def apply(cls, lags, his, his_span, interp_method="cubic"):
        return super().apply(lags, his, his_span, interp_method)

    @staticmethod
    def backward(ctx, grad_y):
        (deriv,) = ctx.saved_tensors
        return history_gather_bwd(grad_y.contiguous(), deriv), None, None, None


class DdeFuse(torch.autograd.Function if torch is not None else object):
    """BaseDDE.fuse (xde/base_dde.py:55-58) with its cotangents: the D3STN trainer backpropagates through the
    solution of ddeint into `func`'s parameters (example/D3STN/train_dde.py:424-454), so the update must stay on
    the autograd graph.  y1 = (dy - 0.001*(dy*dt + y0))*dt + y0."""

    @staticmethod
    def forward(ctx, dy, dt, y0):
        dy_d, y0_d = T.to_dev(dy), T.to_dev(y0)
        if dy_d.shape != y0_d.shape:
            dy_d, y0_d = (a.contiguous() for a in torch.broadcast_tensors(dy_d, y0_d))
        out = torch.empty_like(y0_d)
        check(lib().xde_dde_fuse_f32(T.ptr(dy_d), float(dt), T.ptr(y0_d), y0_d.numel(), T.ptr(out), T.stream()))
        ctx.dt = float(dt)
        ctx.shapes = (tuple(dy.shape), tuple(y0.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        need_dy, need_y0 = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        g_dy = torch.empty_like(g) if need_dy else None
        g_y0 = torch.empty_like(g) if need_y0 else None
        if need_dy or need_y0:
            check(lib().xde_dde_fuse_bwd_f32(T.ptr(g), ctx.dt, g.numel(), T.ptr(g_dy), T.ptr(g_y0), T.stream()))
        if g_dy is not None and g_dy.shape != ctx.shapes[0]:
            g_dy = g_dy.sum_to_size(ctx.shapes[0])
        if g_y0 is not None and g_y0.shape != ctx.shapes[1]:
            g_y0 = g_y0.sum_to_size(ctx.shapes[1])
        return g_dy, None, g_y0


def _as_graph_tensor(x):
    """fp32 CUDA tensor that KEEPS its autograd history (T.to_dev detaches: right for kernel inputs, wrong for
    values the caller differentiates through)."""
    if not isinstance(x, torch.Tensor):
        return T.to_dev(x, like=torch.empty(0, device=T.device()))
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    if not x.is_cuda:
        x = x.to(T.device())
    return x


class BaseDDE(BaseXDE):
    """xde/base_dde.py:14-79: one-shot history resampling, then a fixed-step solve with
    move = func(y_lags, y0) and the damped fuse (lambda = 0.001, :55-58)."""
    kind = "dde"

    def __init__(self, func, y0, t_span, lags, his, his_span, his_processed=False):
        super().__init__(name="DDE", var_nums=1, y0=y0, t_span=t_span)
        self.func = func
        self.lags = lags
        if not his_processed:
            if T.is_torch(lags) and lags.requires_grad:
                self.y_lags = HistoryIndex.apply(lags, his, his_span)
            else:
                self.y_lags = history_gather(lags, his, his_span, "cubic")[0]
        else:
            self.y_lags = his
        self.his, self.his_span = his, his_span
        self.init_y0(y0)

    def move(self, t0, dt, y0):
        return self.func(self.y_lags, y0)

    def fuse(self, dy, dt, y0):
        if torch is not None and (T.is_torch(dy) or T.is_torch(y0)):
            return DdeFuse.apply(_as_graph_tensor(dy), float(dt), _as_graph_tensor(y0))
        dy_d = T.to_dev(dy)
        y0_d = T.to_dev(y0, like=dy_d)
        out = T.empty(tuple(y0_d.shape), y0_d)
        check(lib().xde_dde_fuse_f32(T.ptr(dy_d), float(dt), T.ptr(y0_d), y0_d.numel(), T.ptr(out), T.stream(y0_d)))
        return out
