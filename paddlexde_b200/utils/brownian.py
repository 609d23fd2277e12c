"""Brownian increments from the device-side counter-based generator (include/xde_b200.h:
xde_brownian_increments_f32).  Stands in for the fixed-grid use the SDE solver makes of the reference's
host-side BrownianInterval (paddlexde/utils/brownian/brownian_interval.py:178-240): on a fixed grid only
W(t[n+1]) - W(t[n]) is ever requested."""
from __future__ import annotations

import numpy as np

from .. import _tensor as T
from .._lib import check, lib


def brownian_increments(seed: int, t_span, batch: int, dim: int, offset: int = 0, device="cuda", like=None):
    """dW [len(t)-1, batch, dim], dW[n, b, d] = sqrt(|t[n+1]-t[n]|) * N(0, 1) addressed by
    (n, offset + b, d): the table `sdeint(..., options={"bm_seed": seed, "bm_offset": offset})` uses on
    the fly, bit for bit."""
    if like is None and T.torch is not None:  # a torch tensor on `device` unless `like` names another provider
        like = T.torch.empty(0, device=device)
    t = T.to_dev(np.asarray(T.to_host(t_span), dtype=np.float32), like=like)
    out = T.empty((t.numel() - 1, batch, dim), t)
    check(lib().xde_brownian_increments_f32(int(seed) & (2 ** 64 - 1), int(offset), T.ptr(t), t.numel(), batch, dim,
                                            T.ptr(out), T.stream(t)))
    return out
