from .adaptive_solver import AdaptiveHeun, AdaptiveRKSolver, Bosh3, Dopri5, Dopri8, Fehlberg2  # noqa: F401
from .fixed_solver import RK4, Euler, FixedSolver, Midpoint  # noqa: F401
