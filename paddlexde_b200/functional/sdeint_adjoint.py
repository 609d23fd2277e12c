"""sdeint_adjoint with the reference's signature (paddlexde/functional/sdeint_adjoint.py:213-300).

forward  = sdeint under no_grad (:40-52), one kernel launch, the solution kept at every grid point;
backward = the reference's reverse solve is a placeholder (its `augmented_diffusion` repeats the drift
           dynamics, :136-171, and BaseSDE cannot be instantiated), so what runs here is what it is reaching
           for on the solver's fixed grid: the exact adjoint of the Euler-Maruyama recursion
           (csrc/xde_sde_adj.cu), one kernel launch, parameter gradients summed over the local batch.
Gradient convention of the reference (:229-230): grads for `adjoint_params` only, `y0` receives None."""
from __future__ import annotations

import ctypes as C

from .. import _tensor as T
from .._lib import UnsupportedFieldError, check, lib
from .._tensor import torch  # None when PyTorch is not installed: sde_adjoint_backward still works
from ..field import as_field
from ..solver.adaptive_solver import device_tspan, host_tspan
from ..utils.ode_utils import _rms_norm
from .sdeint import sdeint


def sde_adjoint_backward(drift, diffusion, t_span, y_all, grad_y, *, bm_increments=None, bm_seed=None, bm_offset=0,
                         return_adj_y0=False):
    """SdeintAdjointMethod.backward as a plain function on device buffers.
    y_all, grad_y: [B, T, D] (the fixed solver's layout, every grid point).  -> (g_drift [Pf], g_diffusion [Pg],
    adj_y0 [B, D] | None)."""
    y_d = T.to_dev(y_all, like=grad_y if T.is_torch(grad_y) else None).contiguous()
    g_d = T.to_dev(grad_y, like=y_d).contiguous()
    t_host = host_tspan(t_span)
    Tn, D = t_host.size, drift.d
    if y_d.shape[-1] != D or y_d.shape[-2] != Tn or y_d.shape != g_d.shape:
        raise ValueError("y_all and grad_y must both be [B, T, D] with every grid point stored")
    B = y_d.numel() // (Tn * D)
    gf = T.empty((drift.n_params,), y_d)
    gg = T.empty((diffusion.n_params,), y_d)
    a0 = T.empty((B, D), y_d) if return_adj_y0 else None
    dW = None
    if bm_seed is None:
        dW = T.to_dev(bm_increments, like=y_d)
        if tuple(dW.shape) != (Tn - 1, B, D):
            raise ValueError(f"bm_increments must be [T-1, B, D] = {(Tn - 1, B, D)}, got {tuple(dW.shape)}")
    f, g = drift.c_struct(), diffusion.c_struct()
    t_dev = device_tspan(t_host, y_d)
    check(lib().xde_sde_mlp_adjoint_f32(C.byref(f), C.byref(g), T.ptr(t_dev), Tn, T.ptr(y_d), T.ptr(g_d), B, T.ptr(dW),
                                        0 if bm_seed is None else int(bm_seed) & (2 ** 64 - 1), int(bm_offset),
                                        T.ptr(gf), T.ptr(gg), T.ptr(a0), T.stream(y_d)))
    return gf, gg, a0


class SdeintAdjointMethod(torch.autograd.Function if torch is not None else object):
    @staticmethod
    def forward(ctx, holder, y0, t, *params):
        with torch.no_grad():
            ans = sdeint(holder["drift"], holder["diffusion"], y0, t, holder["solver"], rtol=holder["rtol"],
                         atol=holder["atol"], options=holder["options"])
        ctx.holder, ctx.t = holder, t
        ctx.save_for_backward(ans)
        return ans

    @staticmethod
    def backward(ctx, grad_y):
        h = ctx.holder
        (ans,) = ctx.saved_tensors
        o = h["options"]
        D = h["drift"].d
        B = ans.numel() // (ans.shape[-2] * D)
        gf, gg, _ = sde_adjoint_backward(h["drift"], h["diffusion"], ctx.t, ans.reshape(B, -1, D),
                                         grad_y.contiguous().reshape(B, -1, D), bm_increments=o.get("bm_increments"),
                                         bm_seed=o.get("bm_seed"), bm_offset=o.get("bm_offset", 0))
        if h["allreduce"] is not None:  # 8(e): batch shards sum their parameter gradients
            h["allreduce"](gf)
            h["allreduce"](gg)
        flat = list(h["drift"].grads_like_params(gf)) + list(h["diffusion"].grads_like_params(gg))
        grads = []
        for p, gp in zip(h["params"], flat):
            if isinstance(p, torch.Tensor) and p.requires_grad:
                grads.append(gp.to(p.device).reshape(p.shape).to(p.dtype))
            else:
                grads.append(None)
        return (None, None, None, *grads)  # y0 gets no gradient (functional/sdeint_adjoint.py:229-230)


def sdeint_adjoint(drift, diffusion, y0, t, solver, *, rtol=1e-7, atol=1e-9, options={"norm": _rms_norm},
                   event_fn=None, adjoint_rtol=None, adjoint_atol=None, adjoint_solver=None, adjoint_options=None,
                   adjoint_params=None):
    if torch is None:
        raise ImportError("sdeint_adjoint needs an autograd framework (PyTorch); without one call sdeint(...) and "
                          "sde_adjoint_backward(...) directly")
    f, g = as_field(drift), as_field(diffusion)
    if event_fn is not None:
        raise NotImplementedError("event_fn is not supported (the reference ignores it as well)")
    if adjoint_solver is None:
        adjoint_solver = solver
    if adjoint_solver != solver and options is not None and adjoint_options is None:
        raise ValueError("If `adjoint_method != method` then we cannot infer `adjoint_options` from `options`. So as "
                         "`options` has been passed then `adjoint_options` must be passed as well.")
    if adjoint_solver is not solver:
        raise UnsupportedFieldError("the fused SDE adjoint is the adjoint of the forward scheme itself (Euler)")
    options = dict(options or {})
    options.pop("norm", None)
    if options.get("scheme", "em") != "em":
        raise UnsupportedFieldError("the fused SDE adjoint differentiates the Euler-Maruyama scheme")
    if int(options.get("out_stride", 1)) != 1:
        raise ValueError("sdeint_adjoint needs the forward solution at every grid point (out_stride=1)")
    allreduce = options.pop("grad_allreduce", None)
    params = (tuple(f.parameters()) + tuple(g.parameters())) if adjoint_params is None else tuple(adjoint_params)
    holder = dict(drift=f, diffusion=g, solver=solver, rtol=rtol, atol=atol, options=options, params=params,
                  allreduce=allreduce)
    y0_t = y0 if isinstance(y0, torch.Tensor) else T.to_dev(y0, like=torch.empty(0, device=T.device()))
    if y0_t.dim() != 3 or y0_t.shape[-2] != 1:
        raise ValueError("sdeint_adjoint takes y0 of shape [B, 1, D] (the fixed solvers then return [B, T, D])")
    tensor_params = [p if isinstance(p, torch.Tensor) else torch.as_tensor(p) for p in params]
    return SdeintAdjointMethod.apply(holder, y0_t, t, *tensor_params)
