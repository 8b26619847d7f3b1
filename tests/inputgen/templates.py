"""Automated blocking templates -- host-side restatement of ``src/core/templates/O4H.zig`` (INPUT-GEN, not product code).

The O4H template builds the 8-block cascade topology (two O-grid halves + in/out/down/up/upstream/
downstream), calls ``Block2d.init`` (the TFI hot path) once per block and declares the 21 connections
(3 periodic) and the inlet/outlet conditions (``O4H.zig:423-521``).  It is the caller immediately in
front of the accelerated path and the source of the T106 / LS89 configurations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

from turbomesh_b200 import clustering as cluster
from turbomesh_b200.boundary import Condition, ConditionTag, Connection, Range, Side
from turbomesh_b200.discrete import Block2d, Edge, Mesh

from .edges import EdgeView, combine
from .geometry import Geometry, Line


@dataclass
class NumCells:
    """``O4H.zig:46-65``."""

    o_grid: int
    middle_i: int
    in_up_j: int
    in_down_j: int
    in_i: int
    out_up_j: int
    out_down_j: int
    out_i: int
    down_j: int
    bulge: int
    upstream_i: int
    downstream_i: int

    def scaled(self, factor: int) -> "NumCells":
        return NumCells(**{k: v * factor for k, v in self.__dict__.items()})


def project_normal(edge: np.ndarray, distance: float) -> np.ndarray:
    """``projectNormal``, ``O4H.zig:531-574``: x_i + d * n, n = (t_y, -t_x)/|t| (central, one-sided at the ends)."""
    n = len(edge)
    out = np.empty_like(edge)

    def proj(x_i, tx, ty):
        inv = 1.0 / math.sqrt(tx * tx + ty * ty)
        nx, ny = inv * ty, inv * (-tx)
        return x_i[0] + distance * nx, x_i[1] + distance * ny

    for i in range(1, n - 1):
        tx = 0.5 * (float(edge[i + 1, 0]) - float(edge[i - 1, 0]))
        ty = 0.5 * (float(edge[i + 1, 1]) - float(edge[i - 1, 1]))
        out[i] = proj(edge[i], tx, ty)
    out[0] = proj(edge[0], float(edge[1, 0]) - float(edge[0, 0]), float(edge[1, 1]) - float(edge[0, 1]))
    out[n - 1] = proj(edge[n - 1], float(edge[n - 1, 0]) - float(edge[n - 2, 0]), float(edge[n - 1, 1]) - float(edge[n - 2, 1]))
    return out


@dataclass
class O4H:
    """``O4H.zig:38-65``."""

    blade_clustering: object
    num_cells: NumCells
    inlet_distance: Optional[float] = None
    outlet_distance: Optional[float] = None

    def run(self, geom: Geometry, tfi=None) -> Mesh:
        """``O4H.run``, ``O4H.zig:67-524``."""
        nc = self.num_cells
        uniform = cluster.Uniform()
        num_cells_up = nc.in_up_j + nc.middle_i + nc.bulge + nc.out_up_j + nc.out_i
        num_cells_down = nc.in_down_j + nc.middle_i + nc.out_down_j

        profile_length = geom.profile.up_part.total_length + geom.profile.down_part.total_length
        default_spacing = profile_length / float(num_cells_up + num_cells_down)

        down_edge = Edge.init(num_cells_down + 1, geom.profile.down_part, self.blade_clustering)
        up_edge = Edge.init(num_cells_up + 1, geom.profile.up_part, self.blade_clustering)

        leading_edge = up_edge.points[0].copy()
        down_edge.points[0] = leading_edge
        trailing_edge = up_edge.points[-1].copy()
        down_edge.points[-1] = trailing_edge

        inlet_distance = self.inlet_distance if self.inlet_distance is not None else default_spacing * float(nc.upstream_i)
        outlet_distance = self.outlet_distance if self.outlet_distance is not None else default_spacing * float(nc.downstream_i)

        d = getattr(self, "o_grid_thickness", 0.001)  # O4H.zig:102 hard-codes the O-grid offset 0.001; synthetic passages may widen it
        down_outer_edge = Edge(project_normal(down_edge.points, d), down_edge.clustering.copy())
        up_outer_edge = Edge(project_normal(up_edge.points, -d), up_edge.clustering.copy())
        up_outer_edge.points[0] = down_outer_edge.points[0]
        up_outer_edge.points[-1] = down_outer_edge.points[-1]

        mesh = Mesh()
        pitch = geom.pitch
        le = (float(leading_edge[0]), float(leading_edge[1]))
        te = (float(trailing_edge[0]), float(trailing_edge[1]))

        def pt(a):
            return float(a[0]), float(a[1])

        # Block BLADE_UP (0) -- O4H.zig:115-148
        blade_up_i_min, blade_up_i_max = up_edge, up_outer_edge
        o_cluster = cluster.SingleHyperbolicClustering(delta_s=getattr(self, "o_grid_delta_s", 0.01))  # O4H.zig:129 hard-codes 0.01; refined synthetic passages scale it
        blade_up_j_min = Edge.init(nc.o_grid + 1, Line(pt(blade_up_i_min.points[0]), pt(blade_up_i_max.points[0])), o_cluster)
        blade_up_j_max = Edge.init(nc.o_grid + 1, Line(pt(blade_up_i_min.points[-1]), pt(blade_up_i_max.points[-1])), o_cluster)
        blade_up_id = mesh.add_block("blade_up", Block2d.init(blade_up_i_min, blade_up_i_max, blade_up_j_min, blade_up_j_max, tfi))

        # Block BLADE_DOWN (1) -- O4H.zig:150-166
        blade_down_i_min, blade_down_i_max = down_edge, down_outer_edge
        blade_down_id = mesh.add_block("blade_down", Block2d.init(blade_down_i_min, blade_down_i_max, blade_up_j_min, blade_up_j_max, tfi))

        # Block IN (2) -- O4H.zig:168-212
        in_j_min = combine([EdgeView(blade_up_i_max, nc.in_up_j, 0), EdgeView(blade_down_i_max, 0, nc.in_down_j)])
        assert len(in_j_min.points) == nc.in_up_j + nc.in_down_j + 1
        in_x_00, in_x_01 = pt(in_j_min.points[0]), pt(in_j_min.points[-1])
        in_x_start = le[0] - inlet_distance * 0.5
        in_x_10 = (in_x_start, le[1] + pitch * 0.25)
        in_x_11 = (in_x_start, le[1] - pitch * 0.25)
        in_j_max = Edge.init(len(in_j_min.points), Line(in_x_10, in_x_11), uniform)
        in_i_min = Edge.init(nc.in_i + 1, Line(in_x_00, in_x_10), uniform)
        in_i_max = Edge.init(nc.in_i + 1, Line(in_x_01, in_x_11), uniform)
        in_id = mesh.add_block("in", Block2d.init(in_i_min, in_i_max, in_j_min, in_j_max, tfi))

        # Block OUT (3) -- O4H.zig:214-248
        out_j_min = combine([
            EdgeView(blade_down_i_max, nc.in_down_j + nc.middle_i, len(blade_down_i_max.points) - 1),
            EdgeView(blade_up_i_max, len(blade_up_i_max.points) - 1, nc.in_up_j + nc.bulge + nc.middle_i + nc.out_i),
        ])
        assert len(out_j_min.points) == nc.out_down_j + nc.out_up_j + 1
        out_x_00, out_x_01 = pt(out_j_min.points[0]), pt(out_j_min.points[-1])
        out_x_end = outlet_distance * 0.5 + te[0]
        out_x_10 = (out_x_end, te[1] - pitch * 0.25)
        out_x_11 = (out_x_end, te[1] + pitch * 0.25)
        out_j_max = Edge.init(len(out_j_min.points), Line(out_x_10, out_x_11), uniform)
        out_i_min = Edge.init(nc.out_i + 1, Line(out_x_00, out_x_10), uniform)
        out_i_max = Edge.init(nc.out_i + 1, Line(out_x_01, out_x_11), uniform)
        out_id = mesh.add_block("out", Block2d.init(out_i_min, out_i_max, out_j_min, out_j_max, tfi))

        # Block DOWN (4) -- O4H.zig:250-290
        down_i_min = combine([
            EdgeView(in_i_max, nc.in_i, 0),
            EdgeView(blade_down_i_max, nc.in_down_j, nc.in_down_j + nc.middle_i),
            EdgeView(out_i_min, 0, nc.out_i),
        ])
        down_x_00 = in_x_11
        down_x_01 = (le[0] - 0.0, le[1] - 0.5 * pitch)
        down_x_11 = (te[0] - 0.0, te[1] - 0.5 * pitch)
        down_x_10 = out_x_10
        down_i_max = Edge.init(len(down_i_min.points), Line(down_x_01, down_x_11), uniform)
        down_j_min = Edge.init(nc.down_j + 1, Line(down_x_00, down_x_01), uniform)
        down_j_max = Edge.init(len(down_j_min.points), Line(down_x_10, down_x_11), uniform)
        down_id = mesh.add_block("down", Block2d.init(down_i_min, down_i_max, down_j_min, down_j_max, tfi))

        # Block UP (5) -- O4H.zig:292-346
        up_j_min = out_i_max
        up_i_min = combine([
            EdgeView(blade_up_i_max, nc.in_up_j + nc.middle_i + nc.bulge + nc.out_i, nc.in_up_j),
            EdgeView(in_i_min, 0, nc.in_i),
        ])
        up_x_11 = (le[0] + 0.0, le[1] + 0.5 * pitch)
        up_x_i_max_middle = (te[0] + 0.0, te[1] + 0.5 * pitch)
        up_x_01 = out_x_11
        up_x_10 = in_x_10
        up_i_max_0 = Edge.init(nc.bulge + 1, Line(up_x_01, up_x_i_max_middle), uniform)
        up_i_max_1 = Edge.init(len(up_i_min.points) - nc.bulge, Line(up_x_i_max_middle, up_x_11), uniform)
        up_i_max = combine([EdgeView(up_i_max_0, 0, nc.bulge), EdgeView(up_i_max_1, 0, len(up_i_max_1.points) - 1)])
        up_j_max = Edge.init(nc.out_i + 1, Line(up_x_10, up_x_11), uniform)
        up_id = mesh.add_block("up", Block2d.init(up_i_min, up_i_max, up_j_min, up_j_max, tfi))

        # Block UPSTREAM (6) -- O4H.zig:348-384
        upstream_j_max = combine([
            EdgeView(down_j_min, nc.down_j, 0),
            EdgeView(in_j_max, len(in_j_max.points) - 1, 0),
            EdgeView(up_j_max, 0, len(up_j_max.points) - 1),
        ])
        upstream_x_10, upstream_x_11 = pt(upstream_j_max.points[0]), pt(upstream_j_max.points[-1])
        upstream_x_00 = (le[0] - inlet_distance, le[1] - 0.5 * pitch)
        upstream_x_01 = (le[0] - inlet_distance, le[1] + 0.5 * pitch)
        upstream_j_min = Edge.init(len(upstream_j_max.points), Line(upstream_x_00, upstream_x_01), uniform)
        upstream_i_min = Edge.init(nc.upstream_i + 1, Line(upstream_x_00, upstream_x_10), uniform)
        upstream_i_max = Edge.init(nc.upstream_i + 1, Line(upstream_x_01, upstream_x_11), uniform)
        upstream_id = mesh.add_block("upstream", Block2d.init(upstream_i_min, upstream_i_max, upstream_j_min, upstream_j_max, tfi))

        # Block DOWNSTREAM (7) -- O4H.zig:386-420
        downstream_j_min = combine([
            EdgeView(down_j_max, len(down_j_max.points) - 1, 0),
            EdgeView(out_j_max, 0, len(out_j_max.points) - 1),
            EdgeView(up_i_max_0, 0, len(up_i_max_0.points) - 1),
        ])
        downstream_x_00, downstream_x_01 = pt(downstream_j_min.points[0]), pt(downstream_j_min.points[-1])
        downstream_x_10 = (downstream_x_00[0] + outlet_distance, downstream_x_00[1] + 0.0)
        downstream_x_11 = (downstream_x_10[0] + 0.0, downstream_x_10[1] + pitch)
        downstream_j_max = Edge.init(len(downstream_j_min.points), Line(downstream_x_10, downstream_x_11), uniform)
        downstream_i_min = Edge.init(nc.downstream_i + 1, Line(downstream_x_00, downstream_x_10), uniform)
        downstream_i_max = Edge.init(nc.downstream_i + 1, Line(downstream_x_01, downstream_x_11), uniform)
        downstream_id = mesh.add_block("downstream", Block2d.init(downstream_i_min, downstream_i_max, downstream_j_min, downstream_j_max, tfi))

        # Connections -- O4H.zig:423-515
        S = Side
        n_in_j = len(in_j_min.points)
        n_out_j = len(out_j_max.points)
        n_up_i = len(up_i_min.points)
        n_down_i = len(down_i_min.points)
        n_dsj = len(downstream_j_min.points)

        def conn(b0, s0, a0, e0, b1, s1, a1, e1, per=None):
            mesh.connections.append(Connection((Range(b0, s0, a0, e0), Range(b1, s1, a1, e1)), per))

        conn(blade_up_id, S.j_min, 0, nc.o_grid, blade_down_id, S.j_min, 0, nc.o_grid)
        conn(blade_up_id, S.j_max, 0, nc.o_grid, blade_down_id, S.j_max, 0, nc.o_grid)
        conn(down_id, S.j_min, nc.down_j, 0, upstream_id, S.j_max, 0, nc.down_j)
        conn(in_id, S.j_max, n_in_j - 1, 0, upstream_id, S.j_max, nc.down_j, nc.down_j + n_in_j - 1)
        conn(in_id, S.i_max, 0, nc.in_i, down_id, S.i_min, nc.in_i, 0)
        conn(up_id, S.j_max, 0, nc.out_i, upstream_id, S.j_max, nc.down_j + n_in_j - 1, len(upstream_j_max.points) - 1)
        conn(in_id, S.i_min, 0, nc.in_i, up_id, S.i_min, n_up_i - nc.in_i - 1, n_up_i - 1)
        conn(down_id, S.j_max, nc.down_j, 0, downstream_id, S.j_min, 0, nc.down_j)
        conn(out_id, S.j_max, 0, n_out_j - 1, downstream_id, S.j_min, nc.down_j, nc.down_j + n_out_j - 1)
        conn(out_id, S.i_min, 0, nc.out_i, down_id, S.i_min, n_down_i - 1 - nc.out_i, n_down_i - 1)
        conn(out_id, S.i_max, 0, nc.out_i, up_id, S.j_min, 0, nc.out_i)
        conn(up_id, S.i_max, 0, nc.bulge, downstream_id, S.j_min, n_dsj - 1 - nc.bulge, n_dsj - 1)
        conn(blade_up_id, S.i_max, 0, nc.in_up_j, in_id, S.j_min, nc.in_up_j, 0)
        conn(blade_up_id, S.i_max, nc.in_up_j, nc.in_up_j + nc.middle_i + nc.bulge + nc.out_i, up_id, S.i_min, n_up_i - 1 - nc.in_i, 0)
        conn(blade_up_id, S.i_max, nc.in_up_j + nc.bulge + nc.middle_i + nc.out_i, len(blade_up_i_max.points) - 1, out_id, S.j_min, len(out_j_min.points) - 1, nc.out_down_j)
        conn(blade_down_id, S.i_max, 0, nc.in_down_j, in_id, S.j_min, nc.in_up_j, n_in_j - 1)
        conn(blade_down_id, S.i_max, nc.in_down_j, nc.in_down_j + nc.middle_i, down_id, S.i_min, nc.in_i, n_down_i - 1 - nc.out_i)
        conn(blade_down_id, S.i_max, nc.in_down_j + nc.middle_i, len(blade_down_i_max.points) - 1, out_id, S.j_min, 0, nc.out_down_j)
        per = (0.0, pitch)
        conn(upstream_id, S.i_min, 0, nc.upstream_i, upstream_id, S.i_max, 0, nc.upstream_i, per)
        conn(down_id, S.i_max, 0, len(down_i_max.points) - 1, up_id, S.i_max, len(up_i_max.points) - 1, len(up_i_max.points) - len(down_i_max.points), per)
        conn(downstream_id, S.i_min, 0, nc.downstream_i, downstream_id, S.i_max, 0, nc.downstream_i, per)

        # Boundary conditions -- O4H.zig:518-521
        mesh.boundary_conditions.append(Condition(Range(upstream_id, S.j_min, 0, len(upstream_j_min.points) - 1), ConditionTag.inlet))
        mesh.boundary_conditions.append(Condition(Range(downstream_id, S.j_max, 0, len(downstream_j_max.points) - 1), ConditionTag.outlet))
        return mesh
