from __future__ import annotations

from ..field import as_field
from .base_xde import BaseXDE


class BaseSDE(BaseXDE):
    """paddlexde/xde/base_sde.py:11-61 with repairs R2/R3: tensor state [B, D]; Euler-Maruyama
    ``y1 = y0 + f*dt + g*dW`` (diagonal noise) with the Brownian increments ``dW [T-1, B, D]``
    supplied by the caller instead of the host-side BrownianInterval (xde/base_sde.py:35-37)."""
    kind = "sde"

    def __init__(self, f, g, y0, t_span, reverse=False, bm_increments=None, scheme="em", bm_seed=None, bm_offset=0):
        super().__init__(name="SDE", var_nums=2, y0=y0, t_span=t_span)
        self.f, self.g = f, g
        self.drift, self.diffusion = as_field(f), as_field(g)
        if bm_increments is None and bm_seed is None:
            raise ValueError("sdeint on B200 reads caller-supplied Brownian increments: pass "
                             "options={'bm_increments': dW} with dW of shape [len(t)-1, B, D], or "
                             "options={'bm_seed': int} for the device-side counter-based generator")
        if bm_increments is not None and bm_seed is not None:
            raise ValueError("bm_increments and bm_seed are mutually exclusive")
        self.bm_increments = bm_increments
        # counter-based generator (include/xde_b200.h: xde_sde_mlp_philox_f32): bm_offset = global index of this
        # shard's first trajectory, so that a batch shard sees the increments of the unsharded run
        self.bm_seed, self.bm_offset = bm_seed, int(bm_offset)
        self.scheme = scheme
        self.reverse = reverse
        self.batch_size, self.state_size = y0.shape[0], y0.shape[-1]
        self.init_y0(y0)
