from .adaptive_solver import Dopri5  # noqa: F401
from .fixed_solver import RK4, Euler  # noqa: F401
