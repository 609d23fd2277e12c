from .ode_utils import _mixed_norm, _rms_norm  # noqa: F401
