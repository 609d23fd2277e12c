"""LinearInterpolation / CubicHermiteSpline / BezierSpline with the reference's evaluate/derivative surface
(paddlexde/interpolation/interpolate_base.py:77-114, interpolate.py:6-298).  Nothing is
pre-processed: the gather kernel reads the 2-4 raw neighbours it needs per query."""
from __future__ import annotations

import numpy as np

from .. import _tensor as T
from ..xde.base_dde import history_gather


class _Interp:
    kind = None

    def __init__(self, series, t=None):
        self._series = T.to_dev(series)
        n = self._series.shape[-2]
        if t is None:
            # interpolate_base.py:21-27 builds linspace(0, n, n + 1) but only ever reads its first n entries (one
            # per sample of the series): the grid is 0, 1, ..., n-1
            t = np.arange(n, dtype=np.float32)
        self._t = T.to_dev(t, like=self._series)

    @property
    def grid_points(self):
        return self._t

    def evaluate(self, t):
        return history_gather(self._query(t), self._series, self._t, self.kind)[0]

    def derivative(self, t):
        return history_gather(self._query(t), self._series, self._t, self.kind)[1]

    def _query(self, t):
        if T.is_torch(t):
            return t.reshape(-1)
        return T.to_dev(np.asarray(T.to_host(t), dtype=np.float32).reshape(-1), like=self._series)


class LinearInterpolation(_Interp):
    kind = "linear"


class CubicHermiteSpline(_Interp):
    kind = "cubic"


class BezierSpline(_Interp):
    kind = "bez"
