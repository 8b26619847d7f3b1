"""Curves and blade geometry -- host-side mirror of ``src/core/geometry.zig`` and ``src/core/machine.zig``."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .spline import FittingSpline


@dataclass(frozen=True)
class Line:
    """``geometry.zig:18-40``: ``start + u * (end - start)``."""

    start: tuple
    end: tuple

    def interpolate(self, clustering) -> np.ndarray:
        assert clustering[0] == 0.0 and clustering[-1] == 1.0
        dx = (self.end[0] - self.start[0], self.end[1] - self.start[1])
        out = np.empty((len(clustering), 2), dtype=np.float64)
        for k, u in enumerate(clustering):
            u = float(u)
            out[k, 0] = self.start[0] + u * dx[0]
            out[k, 1] = self.start[1] + u * dx[1]
        return out


class Profile:
    """``machine.zig:17-45``: pressure ("down") and suction ("up") side splines sharing LE and TE."""

    def __init__(self, down, up):
        down = np.asarray(down, dtype=np.float64)
        up = np.asarray(up, dtype=np.float64)
        if not np.array_equal(down[0], up[0]):
            raise ValueError("NonMatchingLeadingEdge")
        if not np.array_equal(down[-1], up[-1]):
            raise ValueError("NonMatchingTrailingEdge")
        assert len(down) > 1 and down[0, 0] < down[-1, 0]
        self.down_part = FittingSpline(down, 3)
        self.up_part = FittingSpline(up, 3)


@dataclass
class Geometry:
    """``machine.zig:8-15``."""

    pitch: float
    profile: Profile
