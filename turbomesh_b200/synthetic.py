"""Synthetic workloads of the named shapes (BASELINE.json configs 3-5).

Everything is deterministic (no RNG).  Generators return a ``discrete.Mesh`` whose blocks carry only their
edges (``EdgeBlock``): the block interiors are produced by the TFI kernel on the device, so even the
512 Mi-node configuration needs only O(boundary) host memory.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .boundary import Condition, ConditionTag, Connection, Range, Side
from .clustering import SingleHyperbolicClustering, Uniform
from .discrete import Edge, Mesh


@dataclass
class EdgeBlock:
    """A block known by its four edges only (input of ``Block2d.init``, discrete.zig:142-159)."""

    i_min: Edge
    i_max: Edge
    j_min: Edge
    j_max: Edge
    points: Optional[np.ndarray] = None

    @property
    def size(self) -> Tuple[int, int]:
        return len(self.i_min.points), len(self.j_min.points)

    def edge_args(self):
        return (self.i_min.points, self.i_max.points, self.j_min.points, self.j_max.points,
                self.i_min.clustering, self.i_max.clustering, self.j_min.clustering, self.j_max.clustering)


def single_block(ni: int = 8192, nj: int = 8192) -> Mesh:
    """Config 3: one block with curved walls, chord c = 1, tanh clustering towards j = 0, no connections.

    i_min (s, 0.08 sin^2(pi s)); i_max (s, 1 + 0.05 sin(2 pi s)); j_min (0.03 sin(pi t), t);
    j_max (1 - 0.03 sin(pi t), t); s uniform; t = Vinokur tanh with delta_s = 0.4/(nj-1).
    """
    s = Uniform().compute(ni)
    t = SingleHyperbolicClustering(0.4 / (nj - 1)).compute(nj)
    sin, pi = np.sin, math.pi
    i_min = np.stack([s, 0.08 * sin(pi * s) ** 2], axis=1)
    i_max = np.stack([s, 1.0 + 0.05 * sin(2 * pi * s)], axis=1)
    j_min = np.stack([0.03 * sin(pi * t), t], axis=1)
    j_max = np.stack([1.0 - 0.03 * sin(pi * t), t], axis=1)
    # exact, shared corner values (tfi.zig:150-162 wants the four corners consistent)
    i_min[0] = j_min[0] = (0.0, 0.0)
    i_min[-1] = j_max[0] = (1.0, 0.0)
    i_max[0] = j_min[-1] = (0.0, 1.0)
    i_max[-1] = j_max[-1] = (1.0, 1.0)
    mesh = Mesh()
    mesh.add_block("block", EdgeBlock(Edge(i_min, s), Edge(i_max, s), Edge(j_min, t), Edge(j_max, t)))
    return mesh


def _channel(xi: np.ndarray, eta: np.ndarray, length: float, height: float, ax: float, ay: float):
    """Wavy channel, periodic in y with period `height`, straight inlet (x = 0) and outlet (x = length)."""
    x = length * xi + ax * np.sin(2 * math.pi * eta) * np.sin(math.pi * xi)
    y = height * eta + ay * np.sin(2 * math.pi * xi)
    return x, y


def cascade(n_bi: int = 8, n_bj: int = 8, ni: int = 4096, nj: int = 2048, length: float = 1.0, height: float = 0.5,
            ax: Optional[float] = None, ay: Optional[float] = None) -> Mesh:
    """Config 4 (tiling form): n_bi x n_bj blocks of ni x nj nodes tiling one cascade passage.

    A wavy channel with inlet at x = 0 and outlet at x = `length`, pitch-wise periodic (period `height`) like the T106
    passage.  The "blade" is a thin plate on the grid line between block rows n_bj//2-1 and n_bj//2, spanning block
    columns n_bi//4 <= bi < n_bi - n_bi//4: there the two block rows are not connected and both faces are fixed walls.
    Exercises interiors, smoothed/connected interfaces, 4-way junctions (also at the plate's leading/trailing edge),
    periodic junctions, walls and sliding nodes.  With n_bj == 1 there is no plate: the channel walls are fixed instead
    of periodic.  Block (bi, bj) has index bi * n_bj + bj, so every connection has ranges[0].block <= ranges[1].block
    (smooth.zig:627).  The domain stays within [0,1] so that the TFI's re-computed boundary nodes of neighbouring blocks
    agree within the reference's 1e-15 absolute coincidence tolerance (smooth.zig:221).
    """
    gi, gj = n_bi * (ni - 1) + 1, n_bj * (nj - 1) + 1
    xi_all = np.arange(gi, dtype=np.float64) / (gi - 1)
    eta_all = np.arange(gj, dtype=np.float64) / (gj - 1)
    # amplitudes of the waviness; the grid lines j = const are inclined by up to atan(2 pi ay / length) against the x axis
    ax = 0.02 * length if ax is None else ax
    ay = 0.03 * height if ay is None else ay
    u_i = Uniform().compute(ni)
    u_j = Uniform().compute(nj)
    mesh = Mesh()

    def line_i(J: int):  # all global nodes of the grid line j = J
        x, y = _channel(xi_all, np.full(gi, eta_all[J]), length, height, ax, ay)
        if J == gj - 1:  # exact periodic image of the line j = 0
            x0, y0 = _channel(xi_all, np.zeros(gi), length, height, ax, ay)
            x, y = x0, y0 + height
        x[0], x[-1] = 0.0, length
        return np.stack([x, y], axis=1)

    def line_j(I: int):
        x, y = _channel(np.full(gj, xi_all[I]), eta_all, length, height, ax, ay)
        if I == 0:
            x[:] = 0.0
        if I == gi - 1:
            x[:] = length
        y0 = _channel(np.array([xi_all[I]]), np.zeros(1), length, height, ax, ay)[1][0]
        y[-1] = y0 + height
        return np.stack([x, y], axis=1)

    lines_i = {bj: line_i(bj * (nj - 1)) for bj in range(n_bj + 1)}
    lines_j = {bi: line_j(bi * (ni - 1)) for bi in range(n_bi + 1)}
    for bi in range(n_bi):
        for bj in range(n_bj):
            I0, J0 = bi * (ni - 1), bj * (nj - 1)
            e_i_min = lines_i[bj][I0:I0 + ni].copy()
            e_i_max = lines_i[bj + 1][I0:I0 + ni].copy()
            e_j_min = lines_j[bi][J0:J0 + nj].copy()
            e_j_max = lines_j[bi + 1][J0:J0 + nj].copy()
            # the i-lines own the corner values
            e_j_min[0], e_j_min[-1] = e_i_min[0], e_i_max[0]
            e_j_max[0], e_j_max[-1] = e_i_min[-1], e_i_max[-1]
            mesh.add_block(f"b{bi}_{bj}", EdgeBlock(Edge(e_i_min, u_i), Edge(e_i_max, u_i), Edge(e_j_min, u_j), Edge(e_j_max, u_j)))

    def idx(bi, bj):
        return bi * n_bj + bj

    S = Side
    plate_row = n_bj // 2  # the plate lies between block rows plate_row-1 and plate_row
    plate_lo, plate_hi = n_bi // 4, n_bi - n_bi // 4
    for bi in range(n_bi):
        for bj in range(n_bj):
            if bi + 1 < n_bi:
                mesh.connections.append(Connection((Range(idx(bi, bj), S.j_max, 0, nj - 1), Range(idx(bi + 1, bj), S.j_min, 0, nj - 1))))
            if bj + 1 < n_bj:
                on_plate = n_bj >= 2 and bj + 1 == plate_row and plate_lo <= bi < plate_hi
                if not on_plate:
                    mesh.connections.append(Connection((Range(idx(bi, bj), S.i_max, 0, ni - 1), Range(idx(bi, bj + 1), S.i_min, 0, ni - 1))))
    if n_bj >= 2:
        for bi in range(n_bi):
            mesh.connections.append(Connection((Range(idx(bi, 0), S.i_min, 0, ni - 1), Range(idx(bi, n_bj - 1), S.i_max, 0, ni - 1)), (0.0, height)))
    for bj in range(n_bj):
        mesh.boundary_conditions.append(Condition(Range(idx(0, bj), S.j_min, 0, nj - 1), ConditionTag.inlet))
        mesh.boundary_conditions.append(Condition(Range(idx(n_bi - 1, bj), S.j_max, 0, nj - 1), ConditionTag.outlet))
    return mesh


def materialize(mesh: Mesh, tfi) -> Mesh:
    """Replaces every ``EdgeBlock`` by a ``Block2d`` using the given TFI callable (GPU or, in tests, the oracle)."""
    from .discrete import Block2d

    out = Mesh(names=list(mesh.names), connections=list(mesh.connections), boundary_conditions=list(mesh.boundary_conditions))
    for b in mesh.blocks:
        out.blocks.append(Block2d(tfi(*b.edge_args())) if isinstance(b, EdgeBlock) else b)
    return out


def batch_of_cuts(base: Mesh, scales) -> Tuple[Mesh, List[Tuple[int, int]]]:
    """Config 5: a batch of independent 2D cuts meshed in one go (the roadmap's 3D-from-2D-cuts path).

    Every cut is the O4H block set of `base` (a mesh of ``EdgeBlock``s, e.g. the T106 inputs) scaled by ``scales[k]``
    (span-wise variation of the blade size; the pitch scales alike).  The cuts are concatenated into ONE mesh with
    block / connection indices offset per cut, so all of them advance in the same kernel launches; there is no
    connection between cuts.  Returns the mesh and the White groups (the two O-grid half blocks of every cut).
    """
    nb = len(base.blocks)
    out = Mesh()
    groups = []
    for k, sc in enumerate(scales):
        sc = float(sc)
        for name, b in zip(base.names, base.blocks):
            def scaled(e: Edge) -> Edge:
                return Edge(e.points * sc, e.clustering.copy())
            out.add_block(f"cut{k}_{name}", EdgeBlock(scaled(b.i_min), scaled(b.i_max), scaled(b.j_min), scaled(b.j_max)))
        for c in base.connections:
            r0, r1 = c.ranges
            per = None if c.periodicity is None else (c.periodicity[0] * sc, c.periodicity[1] * sc)
            out.connections.append(Connection((Range(r0.block + k * nb, r0.side, r0.start, r0.end), Range(r1.block + k * nb, r1.side, r1.start, r1.end)), per))
        for bc in base.boundary_conditions:
            r = bc.range
            out.boundary_conditions.append(Condition(Range(r.block + k * nb, r.side, r.start, r.end), bc.kind))
        groups.append((k * nb, k * nb + 1))
    return out, groups
