"""odeint_adjoint with the reference's signature (paddlexde/functional/odeint_adjoint.py:170-257).

forward  = odeint under no_grad (:38-44)  -> one dopri5 kernel launch
backward = reverse-time solve of the augmented state (y, a, g_theta) segment by segment (:47-167)
           -> one adjoint kernel launch; parameter gradients come back summed over the local batch.
Gradient convention of the reference (:167): grads for `adjoint_params` only, `y0` receives None."""
from __future__ import annotations

import ctypes as C

from .. import _tensor as T
from .._tensor import torch  # None when PyTorch is not installed: adjoint_backward still works (numpy / DeviceArray)
from .._lib import ADJ_NORM, CTRL, UnsupportedFieldError, check, lib, raise_for_status
from ..field import as_field
from ..solver.adaptive_solver import (AttemptLog, Dopri5, StatsBuffer, StatsPair, check_norm, default_controller,
                                      device_tspan, host_tspan, make_ctrl_opts)
from ..utils.ode_utils import _rms_norm
from .odeint import odeint


def adjoint_backward(field, t_span, y_ans, grad_y, *, rtol=1e-7, atol=1e-9, controller="trajectory",
                     adj_norm="seminorm", log_attempts=0, check_status=True, return_adj_y0=False,
                     out_grad_t=None, stats_buffer=None, **ctrl_kw):
    """OdeintAdjointMethod.backward as a plain function on device buffers.

    y_ans, grad_y: [T, B, D] (time-major, as `odeint(..., Dopri5)` returns).  Returns
    (gparams_flat [P], adj_y0 [B, D] | None, stats_reader, attempt_log | None).
    out_grad_t: optional fp32 device tensor [T] that receives grad_t_span (functional/odeint_adjoint.py:129-141,
    161-162; the reference computes it only when t_span requires a gradient)."""
    y_ans_d = T.to_dev(y_ans, like=grad_y if T.is_torch(grad_y) else None)
    grad_d = T.to_dev(grad_y, like=y_ans_d)
    t_host = host_tspan(t_span)
    Tn = t_host.size
    D = field.d
    if y_ans_d.shape[0] != Tn or y_ans_d.shape[-1] != D or y_ans_d.shape != grad_d.shape:
        raise ValueError("y_ans and grad_y must both be [T, ..., D]")
    B = y_ans_d.numel() // (Tn * D)
    g = T.zeros((field.n_params,), y_ans_d)
    a0 = T.empty((B, D), y_ans_d) if return_adj_y0 else None
    stats = StatsBuffer(y_ans_d) if stats_buffer is None else stats_buffer
    log = AttemptLog(B, log_attempts, y_ans_d) if log_attempts > 0 else None
    opts = make_ctrl_opts(rtol, atol, **ctrl_kw)
    fs = field.c_struct()
    t_dev = device_tspan(t_host, y_ans_d)
    check(lib().xde_dopri5_mlp_adjoint_f32(C.byref(fs), T.ptr(t_dev), Tn, T.ptr(y_ans_d), T.ptr(grad_d), B,
                                           C.byref(opts), CTRL[controller], ADJ_NORM[adj_norm], T.ptr(g),
                                           T.ptr(a0), T.ptr(out_grad_t), T.ptr(stats.buf),
                                           C.byref(log.c_struct()) if log else None, T.stream(y_ans_d)))
    if check_status:
        raise_for_status(stats.read().status)
    return g, a0, stats, log


class OdeintAdjointMethod(torch.autograd.Function if torch is not None else object):
    """The autograd surface needs an autograd framework: this is the PyTorch adapter (INTEGRATION.md shows the Paddle
    `PyLayer` twin).  Without PyTorch installed the class is inert and `odeint_adjoint` raises ImportError; the
    numerical work is `adjoint_backward` above either way."""

    @staticmethod
    def forward(ctx, holder, y0, t_span, *params):
        with torch.no_grad():
            ans = odeint(holder["field"], y0, t_span, holder["solver"], rtol=holder["rtol"], atol=holder["atol"],
                         options=holder["options"])
        holder["fwd_solver"] = odeint.last_solver
        ctx.holder = holder
        ctx.t_span = t_span
        ctx.save_for_backward(ans)
        return ans

    @staticmethod
    def backward(ctx, grad_y):
        h = ctx.holder
        (ans,) = ctx.saved_tensors
        if h["adjoint_solver"] is not Dopri5:
            raise NotImplementedError("the fused adjoint backward integrates with Dopri5")
        defer = h.get("defer_fwd_status", False)
        # grad_t_span only when t_span requires a gradient (t_requires_grad, functional/odeint_adjoint.py:27,130-141)
        t_req = isinstance(ctx.t_span, torch.Tensor) and ctx.needs_input_grad[2]
        grad_t = torch.empty(ctx.t_span.numel(), device=ans.device, dtype=torch.float32) if t_req else None
        g, _, stats, _ = adjoint_backward(h["field"], ctx.t_span, ans, grad_y.contiguous(), rtol=h["adjoint_rtol"],
                                          atol=h["adjoint_atol"], controller=h["controller"],
                                          adj_norm=h["adj_norm"], check_status=False, out_grad_t=grad_t,
                                          stats_buffer=h["stats_pair"].adj, **h["adjoint_ctrl"])
        h["bwd_stats"] = stats
        # both solves are queued; ONE device-to-host copy brings both status words.  The forward assertion first
        # (it is the cause) when its check was deferred to here, then the adjoint's.
        if h.get("check_status", True) is not False:
            st_fwd, st_adj = h["stats_pair"].read()
            if defer:
                h["fwd_solver"].stats = st_fwd
                raise_for_status(st_fwd.status)
            raise_for_status(st_adj.status)
        # check_status=False: nothing is read back (the call sequence stays capturable in a CUDA graph); the caller
        # reads `odeint_adjoint.last["stats_pair"]` when it wants the status words
        if h["allreduce"] is not None:
            h["allreduce"](g)  # 8(e): the only collective on the path (adjoint parameter gradients)
        grads = []
        for p, gp in zip(h["params"], h["field"].grads_like_params(g)):
            if isinstance(p, torch.Tensor) and p.requires_grad:
                grads.append(gp.to(p.device).reshape(p.shape).to(p.dtype))
            else:
                grads.append(None)
        if grad_t is not None:
            grad_t = grad_t.to(ctx.t_span.device).reshape(ctx.t_span.shape).to(ctx.t_span.dtype)
        return (None, None, grad_t, *grads)  # y0 gets no gradient (functional/odeint_adjoint.py:167)


def odeint_adjoint(func, y0, t_span, *, rtol=1e-7, atol=1e-9, solver=None, options={"norm": _rms_norm},
                   event_fn=None, adjoint_rtol=None, adjoint_atol=None, adjoint_solver=None,
                   adjoint_options=None, adjoint_params=None):
    if torch is None:
        raise ImportError("odeint_adjoint returns a tensor with an autograd graph: it needs PyTorch (or the Paddle PyLayer "
                          "of INTEGRATION.md). Without an autograd framework call odeint(...) and "
                          "paddlexde_b200.functional.odeint_adjoint.adjoint_backward(...) directly.")
    field = as_field(func)  # nn.Layer check of the reference (:186-192) becomes: must be a fused field
    if event_fn is not None:
        raise NotImplementedError("event_fn is not supported (the reference ignores it as well)")
    # option defaulting exactly as the reference (:194-214)
    if adjoint_rtol is None:
        adjoint_rtol = rtol
    if adjoint_atol is None:
        adjoint_atol = atol
    if adjoint_solver is None:
        adjoint_solver = solver
    if adjoint_solver != solver and options is not None and adjoint_options is None:
        raise ValueError("If `adjoint_method != method` then we cannot infer `adjoint_options` from `options`. So as "
                         "`options` has been passed then `adjoint_options` must be passed as well.")
    if adjoint_options is None:
        adjoint_options = {k: v for k, v in options.items() if k != "norm"} if options is not None else {}
    else:
        adjoint_options = adjoint_options.copy()
    params = tuple(field.parameters()) if adjoint_params is None else tuple(adjoint_params)
    if options is not None:
        check_norm(options.get("norm"))

    # adjoint norm (handle_adjoint_norm_, :280-327): default mixed norm | "seminorm"
    controller = adjoint_options.pop("controller", (options or {}).get("controller"))
    if controller is None:  # announced once: the reference's default is the batch controller + mixed norm
        shp = getattr(y0, "shape", None)
        rows = None
        if shp is not None and len(shp) >= 1:
            rows = 1
            for v in shp[:-1]:
                rows *= int(v)
        controller = default_controller(rows)
        options = {**(options or {}), "controller": controller}
    norm = adjoint_options.pop("norm", None)
    if norm is None:
        # With one controller per trajectory the parameter-gradient state is a per-trajectory partial
        # integral, so the reference's default mixed norm is only meaningful for controller="batch"
        # (SURVEY 7.3.1).  The per-trajectory controller therefore uses the seminorm.
        adj_norm = "mixed" if controller == "batch" else "seminorm"
    elif norm == "seminorm":
        adj_norm = "seminorm"
    else:
        raise UnsupportedFieldError("custom adjoint norm callables cannot be fused; use 'seminorm' or the default")
    adjoint_ctrl = {k: adjoint_options[k] for k in ("min_step", "max_step", "first_step", "safety", "ifactor",
                                                    "dfactor", "max_num_steps") if k in adjoint_options}
    # When the forward solve's assertions (dt underflow, non-finite state, max_num_steps: base_adaptive_solver_rk.py:
    # 120-122, 200-203) are raised:
    #   options={"check_status": True}        by the call itself, like the reference (one host round trip per solve);
    #   options={"check_status": "deferred"}  by backward(), after the adjoint solve has been queued -- the GPU does
    #                                         not idle while the host walks from one solve to the other;
    #   options={"check_status": False}       never (the caller reads `odeint_adjoint.last["fwd_solver"].read_stats()`).
    # Default: "deferred" when some adjoint parameter requires a gradient (a training step: backward() follows),
    # True otherwise.
    cs = (options or {}).get("check_status", None)
    if cs is None:
        training = any(isinstance(p, torch.Tensor) and p.requires_grad for p in params)
        cs = "deferred" if training else True
    defer = (cs == "deferred")
    options = {**(options or {}), "check_status": False if defer else cs}
    holder = dict(field=field, solver=solver, rtol=rtol, atol=atol, options=options or {}, defer_fwd_status=defer,
                  check_status=cs,
                  adjoint_rtol=adjoint_rtol, adjoint_atol=adjoint_atol, adjoint_solver=adjoint_solver,
                  adjoint_ctrl=adjoint_ctrl, controller=controller, adj_norm=adj_norm, params=params,
                  allreduce=(options or {}).get("grad_allreduce"))
    if "grad_allreduce" in holder["options"]:
        holder["options"] = {k: v for k, v in holder["options"].items() if k != "grad_allreduce"}
    y0_t = y0 if isinstance(y0, torch.Tensor) else T.to_dev(y0, like=torch.empty(0, device=T.device()))
    if not y0_t.is_cuda:
        y0_t = y0_t.to(T.device())
    holder["stats_pair"] = StatsPair(y0_t)
    holder["options"]["stats_buffer"] = holder["stats_pair"].fwd
    tensor_params = [p if isinstance(p, torch.Tensor) else torch.as_tensor(p) for p in params]
    sol = OdeintAdjointMethod.apply(holder, y0_t, t_span, *tensor_params)
    odeint_adjoint.last = holder
    return sol
