from .interpolate import CubicHermiteSpline, LinearInterpolation  # noqa: F401
