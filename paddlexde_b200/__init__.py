"""paddlexde_b200 -- B200-native (sm_100a) drop-in for PaddleXDE's batched DE-integration hot path.

Mirrors the reference's import surface (paddlexde/__init__.py:4-8): the functional entry points at
top level, solver classes under .solver, problem wrappers under .xde, interpolants under
.interpolation.  Everything numerical runs in hand-written CUDA kernels behind the C ABI of
include/xde_b200.h; there is no CPU fallback."""
from ._lib import UnsupportedFieldError, XdeError, launch_count, library_path  # noqa: F401
from .field import MLPField  # noqa: F401
from .functional import ddeint, ddeint_adjoint, odeint, odeint_adjoint, sdeint, sdeint_adjoint  # noqa: F401
from .solver import RK4, AdaptiveHeun, Bosh3, Dopri5, Dopri8, Euler, Fehlberg2, Midpoint  # noqa: F401
from . import interpolation, solver, utils, xde  # noqa: F401

__version__ = "0.1.0"
