"""``Edge.combine`` and ``EdgeView`` -- host-side restatement of ``src/core/discrete.zig:38-136`` (INPUT-GEN: used to reproduce the
reference's example inputs for fixtures and workloads; not part of the product package)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

from turbomesh_b200.discrete import Edge


def combine(edges: Sequence["EdgeView"]) -> Edge:
    """``Edge.combine``, ``discrete.zig:38-91``: concatenates views, dropping the duplicated joints."""
    assert len(edges) > 1
    tol = 1e-10
    for k in range(len(edges) - 1):
        a = edges[k].edge.points[edges[k].end]
        b = edges[k + 1].edge.points[edges[k + 1].start]
        if not (abs(a[0] - b[0]) <= tol and abs(a[1] - b[1]) <= tol):
            raise ValueError(f"edges {k + 1} and {k + 2} cannot be combined as end points do not match: {a} and {b}")
    n = sum(e.len() for e in edges) - (len(edges) - 1)
    u = np.empty(n, dtype=np.float64)
    points = np.empty((n, 2), dtype=np.float64)
    start = 0
    for e in edges:
        start += e.clone_points(points[start:]) - 1
    start = 0
    last_value = 0.0
    for e in edges:
        start += e.clone_clustering(u[start:], last_value) - 1
        last_value = float(u[start])
    for k in range(n):
        u[k] = u[k] / last_value
    return Edge(points, u)


@dataclass
class EdgeView:
    """``discrete.zig:94-136``."""

    edge: Edge
    start: int
    end: int

    def len(self) -> int:
        return abs(self.start - self.end) + 1

    def clone_points(self, buffer: np.ndarray) -> int:
        n = self.len()
        if self.start > self.end:
            buffer[:n] = self.edge.points[self.end : self.start + 1][::-1]
        else:
            buffer[:n] = self.edge.points[self.start : self.end + 1]
        return n

    def clone_clustering(self, buffer: np.ndarray, initial_value: float) -> int:
        # NOTE (reference behaviour): for reversed views the deltas are still accumulated from
        # min(start, end) upwards, i.e. the clustering is not reversed (discrete.zig:119-135).
        buffer[0] = initial_value
        first, last = min(self.start, self.end), max(self.start, self.end)
        last_value = float(self.edge.clustering[first])
        i_buf = 1
        for i in range(first + 1, last + 1):
            delta = float(self.edge.clustering[i]) - last_value
            buffer[i_buf] = initial_value + delta
            i_buf += 1
        return i_buf
