from . import brownian  # noqa: F401  (px.utils.brownian.brownian_increments is part of the public surface)
from .ode_utils import _linf_norm, _mixed_norm, _rms_norm  # noqa: F401
