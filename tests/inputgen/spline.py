"""Natural cubic fitting spline -- host-side mirror of ``src/core/spline.zig:10-233``.

Chord-length parameterised natural cubic spline with a 201-sample arc-length lookup table;
``interpolate(u)`` maps normalised arc length to points.  Used to discretise blade edges
(INPUT-GEN for the TFI + smoothing path).
"""
from __future__ import annotations

import math

import numpy as np

SAMPLE_COUNT = 200  # spline.zig:22


def _distance(a, b) -> float:
    s = 0.0
    for k in range(len(a)):
        d = b[k] - a[k]
        s += d * d
    return math.sqrt(s)


class FittingSpline:
    def __init__(self, points, degree: int = 3):
        if degree != 3:
            raise ValueError("UnsupportedDegree")
        pts = np.asarray(points, dtype=np.float64)
        if pts.ndim != 2 or pts.shape[0] < 2:
            raise ValueError("NotEnoughPoints")
        self.points = pts.copy()
        self.dim = pts.shape[1]
        n = pts.shape[0]
        # computeChordParams, spline.zig:141-155
        params = [0.0] * n
        total = 0.0
        for i in range(1, n):
            total += _distance(pts[i - 1], pts[i])
            params[i] = total
        if total == 0.0:
            params = [i / (n - 1) for i in range(n)]
        else:
            params = [p / total for p in params]
        self.params = params
        self.total_length = total
        self.second_derivs = [self._second_derivs(d) for d in range(self.dim)]
        self._build_arc_length_table()

    def _second_derivs(self, comp: int):
        # computeSecondDerivs, spline.zig:157-200
        params, pts = self.params, self.points
        n = len(params)
        z = [0.0] * n
        if n == 2:
            return z
        tmp = [0.0] * n
        for i in range(1, n - 1):
            h_im1 = params[i] - params[i - 1]
            h_i = params[i + 1] - params[i]
            if h_im1 == 0.0 or h_i == 0.0:
                raise ValueError("CoincidentParameters")
            dy_im1 = float(pts[i][comp]) - float(pts[i - 1][comp])
            dy_i = float(pts[i + 1][comp]) - float(pts[i][comp])
            alpha = (dy_i / h_i) - (dy_im1 / h_im1)
            denom = 2.0 * (params[i + 1] - params[i - 1]) - h_im1 * tmp[i - 1]
            tmp[i] = h_i / denom
            z[i] = (6.0 * alpha - h_im1 * z[i - 1]) / denom
        z[n - 1] = 0.0
        for k in range(n - 2, -1, -1):
            z[k] = z[k] - tmp[k] * z[k + 1]
        return z

    def eval(self, param: float):
        # spline.zig:202-222
        params = self.params
        u = min(max(param, 0.0), 1.0)
        idx = 0
        while idx + 1 < len(params) and params[idx + 1] < u:
            idx += 1
        if idx >= len(params) - 1:
            idx = len(params) - 2
        h = params[idx + 1] - params[idx]
        a = (params[idx + 1] - u) / h
        b = (u - params[idx]) / h
        out = [0.0] * self.dim
        for d in range(self.dim):
            y0 = float(self.points[idx][d])
            y1 = float(self.points[idx + 1][d])
            z0 = self.second_derivs[d][idx]
            z1 = self.second_derivs[d][idx + 1]
            out[d] = a * y0 + b * y1 + ((a * a * a - a) * z0 + (b * b * b - b) * z1) * (h * h) / 6.0
        return out

    def _build_arc_length_table(self):
        # spline.zig:87-110
        m = SAMPLE_COUNT + 1
        self.sample_params = [i / (m - 1) for i in range(m)]
        self.sample_arc = [0.0] * m
        length = 0.0
        prev = self.eval(self.sample_params[0])
        for k in range(1, m):
            cur = self.eval(self.sample_params[k])
            length += _distance(prev, cur)
            self.sample_arc[k] = length
            prev = cur
        self.total_length = length
        if length == 0.0:
            self.sample_arc = [0.0] * m
            return
        self.sample_arc = [v / length for v in self.sample_arc]

    def param_at_arc_fraction(self, u: float) -> float:
        # spline.zig:112-139
        if self.total_length == 0.0:
            return 0.0
        target = min(max(u, 0.0), 1.0)
        arc = self.sample_arc
        lo, hi = 0, len(arc) - 1
        while lo < hi:
            mid = (lo + hi) // 2
            if arc[mid] < target:
                lo = mid + 1
            else:
                hi = mid
        if lo == 0:
            return self.sample_params[0]
        a0, a1 = arc[lo - 1], arc[lo]
        p0, p1 = self.sample_params[lo - 1], self.sample_params[lo]
        t = (target - a0) / (a1 - a0) if a1 > a0 else 0.0
        return p0 + t * (p1 - p0)

    def interpolate(self, u) -> np.ndarray:
        """``FittingSpline.interpolate``, ``spline.zig:74-81``."""
        out = np.empty((len(u), self.dim), dtype=np.float64)
        for k, uv in enumerate(u):
            out[k] = self.eval(self.param_at_arc_fraction(float(uv)))
        return out

    def integrate(self) -> float:
        return self.total_length
