from .interpolate import BezierSpline, CubicHermiteSpline, LinearInterpolation  # noqa: F401
