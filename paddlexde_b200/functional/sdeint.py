from __future__ import annotations

from ..utils.ode_utils import _rms_norm
from ..xde import BaseSDE


def sdeint(drift, diffusion, y0, t, solver, *, rtol=1e-7, atol=1e-9, reverse=False,
           options: object = {"norm": _rms_norm}):
    """Same signature as paddlexde/functional/sdeint.py:9-37.  Extension (repair R3): the Brownian
    increments come from the caller as options["bm_increments"] = dW [len(t)-1, B, D] (parity mode), or
    from the device-side counter-based generator with options["bm_seed"] (+ "bm_offset" for batch shards);
    options["scheme"] in {"em", "milstein"} ("milstein" has no reference counterpart)."""
    options = dict(options)
    dW = options.pop("bm_increments", None)
    scheme = options.pop("scheme", "em")
    xde = BaseSDE(f=drift, g=diffusion, y0=y0, t_span=t, reverse=reverse, bm_increments=dW, scheme=scheme,
                  bm_seed=options.pop("bm_seed", None), bm_offset=options.pop("bm_offset", 0))
    s = solver(xde=xde, y0=xde.y0, rtol=rtol, atol=atol, **options)
    solution = s.integrate(t)
    return xde.format(solution)
