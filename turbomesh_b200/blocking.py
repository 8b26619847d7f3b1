"""Automated O4H blocking of a batch of 2D cuts on the GPU -- the caller right in front of the accelerated path
(what ``templates.O4H.run`` does for one geometry, ``src/core/templates/O4H.zig:67-524``).

The reference builds one passage edge by edge in program order.  Here the template is DATA -- tables of corner points,
straight edges, combined edges, blocks and connections in terms of the cell counts -- and a batch of cuts goes through it
in six device launches whatever its size:

    1. ``tm_splines_fit``            both blade sides of every cut          (spline.zig:41-110)
    2. ``tm_edges_discretize``       the blade edges                        (discrete.zig:17-31)
    3. ``tm_edges_project_normal``   the outer lines of the O-grid          (O4H.zig:531-574)
    4. ``tm_edges_discretize``       the 14 straight edges whose end points are known by now
    5. ``tm_edges_combine``          the 7 combined edges                   (discrete.zig:38-91)
    6. ``tm_edges_discretize``       the 6 straight edges of the upstream / downstream blocks, which start at end points of 5

The result is a mesh of ``EdgeBlock``s (8 per cut, the cuts concatenated, no connection between cuts) for
``tm_mesh_tfi_block`` + ``tm_mesh_smooth``, and the White groups.  Every edge is bit-identical to what the sequential host
restatement of ``O4H.run`` kept with the tests (the checker) produces: the kernels reproduce the reference's operation
order and the few host operations here (corner points from the leading / trailing edge, pitch and distances) are written
exactly as ``O4H.zig`` writes them.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import clustering as cluster
from .boundary import Condition, ConditionTag, Connection, Range, Side
from .discrete import Edge, FittedSpline, Mesh
from .synthetic import EdgeBlock

BLOCKS = ("blade_up", "blade_down", "in", "out", "down", "up", "upstream", "downstream")   # block order of O4H.zig:115-420


@dataclass(frozen=True)
class Cells:
    """Cell counts of the template, ``O4H.zig:46-65``."""

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

    @property
    def up(self) -> int:        # cells along the upper blade side, O4H.zig:74-76
        return self.in_up_j + self.middle_i + self.bulge + self.out_up_j + self.out_i

    @property
    def down(self) -> int:
        return self.in_down_j + self.middle_i + self.out_down_j


class _Line:
    def __init__(self, start, end):
        self.start, self.end = (float(start[0]), float(start[1])), (float(end[0]), float(end[1]))


@lru_cache(maxsize=None)
def _views(c: Cells) -> Dict[str, List[Tuple[str, int, int]]]:
    """The combined edges as lists of (edge, first index, last index) -- ``EdgeView``s, O4H.zig:168-420."""
    nu, nd = c.up, c.down
    n_in_j = c.in_up_j + c.in_down_j + 1
    n_out_j = c.out_down_j + c.out_up_j + 1
    n_up_i = (c.in_up_j + c.middle_i + c.bulge + c.out_i - c.in_up_j) + c.in_i + 1
    return {
        "in_j_min": [("up_outer", c.in_up_j, 0), ("down_outer", 0, c.in_down_j)],
        "out_j_min": [("down_outer", c.in_down_j + c.middle_i, nd), ("up_outer", nu, c.in_up_j + c.bulge + c.middle_i + c.out_i)],
        "down_i_min": [("in_i_max", c.in_i, 0), ("down_outer", c.in_down_j, c.in_down_j + c.middle_i), ("out_i_min", 0, c.out_i)],
        "up_i_min": [("up_outer", c.in_up_j + c.middle_i + c.bulge + c.out_i, c.in_up_j), ("in_i_min", 0, c.in_i)],
        "up_i_max": [("up_i_max_0", 0, c.bulge), ("up_i_max_1", 0, n_up_i - c.bulge - 1)],
        "upstream_j_max": [("down_j_min", c.down_j, 0), ("in_j_max", n_in_j - 1, 0), ("up_j_max", 0, c.out_i)],
        "downstream_j_min": [("down_j_max", c.down_j, 0), ("out_j_max", 0, n_out_j - 1), ("up_i_max_0", 0, c.bulge)],
    }


@lru_cache(maxsize=None)
def _lengths(c: Cells) -> Dict[str, int]:
    """Number of points of every edge, from the cell counts alone."""
    v = _views(c)
    n = {"up_edge": c.up + 1, "down_edge": c.down + 1, "up_outer": c.up + 1, "down_outer": c.down + 1}
    n_in_j = c.in_up_j + c.in_down_j + 1
    n_out_j = c.out_down_j + c.out_up_j + 1
    n_up_i = sum(abs(a - b) for _, a, b in v["up_i_min"]) + 1
    n_down_i = sum(abs(a - b) for _, a, b in v["down_i_min"]) + 1
    n.update({"o_j_min": c.o_grid + 1, "o_j_max": c.o_grid + 1,
              "in_j_max": n_in_j, "in_i_min": c.in_i + 1, "in_i_max": c.in_i + 1,
              "out_j_max": n_out_j, "out_i_min": c.out_i + 1, "out_i_max": c.out_i + 1,
              "down_i_max": n_down_i, "down_j_min": c.down_j + 1, "down_j_max": c.down_j + 1,
              "up_i_max_0": c.bulge + 1, "up_i_max_1": n_up_i - c.bulge, "up_j_max": c.out_i + 1,
              "upstream_i_min": c.upstream_i + 1, "upstream_i_max": c.upstream_i + 1,
              "downstream_i_min": c.downstream_i + 1, "downstream_i_max": c.downstream_i + 1})
    for name, views in v.items():
        n[name] = sum(abs(a - b) for _, a, b in views) + 1
    n["upstream_j_min"] = n["upstream_j_max"]
    n["downstream_j_max"] = n["downstream_j_min"]
    return n


# blocks as (i_min, i_max, j_min, j_max), O4H.zig:115-420
_BLOCK_EDGES = {
    "blade_up": ("up_edge", "up_outer", "o_j_min", "o_j_max"),
    "blade_down": ("down_edge", "down_outer", "o_j_min", "o_j_max"),
    "in": ("in_i_min", "in_i_max", "in_j_min", "in_j_max"),
    "out": ("out_i_min", "out_i_max", "out_j_min", "out_j_max"),
    "down": ("down_i_min", "down_i_max", "down_j_min", "down_j_max"),
    "up": ("up_i_min", "up_i_max", "out_i_max", "up_j_max"),
    "upstream": ("upstream_i_min", "upstream_i_max", "upstream_j_min", "upstream_j_max"),
    "downstream": ("downstream_i_min", "downstream_i_max", "downstream_j_min", "downstream_j_max"),
}


def connections(c: Cells, pitch: float, first_block: int = 0) -> Tuple[List[Connection], List[Condition]]:
    """The 21 connections (3 periodic) and the inlet / outlet conditions of one passage, ``O4H.zig:423-521``."""
    n = _lengths(c)
    B = {name: first_block + k for k, name in enumerate(BLOCKS)}
    S = Side
    last = lambda e: n[e] - 1  # noqa: E731
    table = [
        ("blade_up", S.j_min, 0, c.o_grid, "blade_down", S.j_min, 0, c.o_grid, False),
        ("blade_up", S.j_max, 0, c.o_grid, "blade_down", S.j_max, 0, c.o_grid, False),
        ("down", S.j_min, c.down_j, 0, "upstream", S.j_max, 0, c.down_j, False),
        ("in", S.j_max, last("in_j_max"), 0, "upstream", S.j_max, c.down_j, c.down_j + last("in_j_max"), False),
        ("in", S.i_max, 0, c.in_i, "down", S.i_min, c.in_i, 0, False),
        ("up", S.j_max, 0, c.out_i, "upstream", S.j_max, c.down_j + last("in_j_max"), last("upstream_j_max"), False),
        ("in", S.i_min, 0, c.in_i, "up", S.i_min, last("up_i_min") - c.in_i, last("up_i_min"), False),
        ("down", S.j_max, c.down_j, 0, "downstream", S.j_min, 0, c.down_j, False),
        ("out", S.j_max, 0, last("out_j_max"), "downstream", S.j_min, c.down_j, c.down_j + last("out_j_max"), False),
        ("out", S.i_min, 0, c.out_i, "down", S.i_min, last("down_i_min") - c.out_i, last("down_i_min"), False),
        ("out", S.i_max, 0, c.out_i, "up", S.j_min, 0, c.out_i, False),
        ("up", S.i_max, 0, c.bulge, "downstream", S.j_min, last("downstream_j_min") - c.bulge, last("downstream_j_min"), False),
        ("blade_up", S.i_max, 0, c.in_up_j, "in", S.j_min, c.in_up_j, 0, False),
        ("blade_up", S.i_max, c.in_up_j, c.in_up_j + c.middle_i + c.bulge + c.out_i, "up", S.i_min, last("up_i_min") - c.in_i, 0, False),
        ("blade_up", S.i_max, c.in_up_j + c.bulge + c.middle_i + c.out_i, last("up_outer"), "out", S.j_min, last("out_j_min"), c.out_down_j, False),
        ("blade_down", S.i_max, 0, c.in_down_j, "in", S.j_min, c.in_up_j, last("in_j_min"), False),
        ("blade_down", S.i_max, c.in_down_j, c.in_down_j + c.middle_i, "down", S.i_min, c.in_i, last("down_i_min") - c.out_i, False),
        ("blade_down", S.i_max, c.in_down_j + c.middle_i, last("down_outer"), "out", S.j_min, 0, c.out_down_j, False),
        ("upstream", S.i_min, 0, c.upstream_i, "upstream", S.i_max, 0, c.upstream_i, True),
        ("down", S.i_max, 0, last("down_i_max"), "up", S.i_max, last("up_i_max"), last("up_i_max") - last("down_i_max"), True),
        ("downstream", S.i_min, 0, c.downstream_i, "downstream", S.i_max, 0, c.downstream_i, True),
    ]
    conns = [Connection((Range(B[b0], s0, a0, e0), Range(B[b1], s1, a1, e1)), (0.0, float(pitch)) if periodic else None)
             for b0, s0, a0, e0, b1, s1, a1, e1, periodic in table]
    conds = [Condition(Range(B["upstream"], S.j_min, 0, last("upstream_j_min")), ConditionTag.inlet),
             Condition(Range(B["downstream"], S.j_max, 0, last("downstream_j_max")), ConditionTag.outlet)]
    return conns, conds


@dataclass
class Cut:
    """One 2D cut: the points the two blade sides are fitted through (leading to trailing edge) and the pitch."""

    profile_up: np.ndarray
    profile_down: np.ndarray
    pitch: float


class O4HBatch:
    """The O4H template for a batch of cuts; parameters as ``templates.O4H`` (``O4H.zig:38-65``; the O-grid offset 0.001 and
    the wall spacing 0.01 of its j edges are hard-coded there, ``O4H.zig:102, 129``)."""

    def __init__(self, cells: Cells, blade_clustering, inlet_distance: Optional[float] = None, outlet_distance: Optional[float] = None,
                 o_grid_thickness: float = 0.001, o_grid_delta_s: float = 0.01, device: int = -1):
        self.cells, self.blade_clustering = cells, blade_clustering
        self.inlet_distance, self.outlet_distance = inlet_distance, outlet_distance
        self.o_grid_thickness, self.o_grid_delta_s, self.device = o_grid_thickness, o_grid_delta_s, device

    def run(self, cuts: Sequence[Cut]) -> Tuple[Mesh, List[Tuple[int, int]]]:
        c, dev = self.cells, self.device
        n = _lengths(c)
        views = _views(c)
        uniform = cluster.Uniform()
        o_cluster = cluster.SingleHyperbolicClustering(delta_s=self.o_grid_delta_s)
        E: List[Dict[str, Edge]] = [dict() for _ in cuts]

        # 1, 2: blade sides
        splines = FittedSpline.fit_batch([p for cut in cuts for p in (cut.profile_up, cut.profile_down)], device=dev)
        blade = Edge.init_batch([(n[name], splines[2 * k + s], self.blade_clustering) for k in range(len(cuts)) for s, name in enumerate(("up_edge", "down_edge"))], dev)
        for k in range(len(cuts)):
            up, down = blade[2 * k], blade[2 * k + 1]
            down.points[0] = up.points[0]          # one leading / trailing edge for both sides, O4H.zig:88-92
            down.points[-1] = up.points[-1]
            E[k]["up_edge"], E[k]["down_edge"] = up, down
        # 3: O-grid outer lines
        d = self.o_grid_thickness
        outer = Edge.project_normal_batch([job for k in range(len(cuts)) for job in ((E[k]["down_edge"].points, d), (E[k]["up_edge"].points, -d))], dev)
        corner: List[Dict[str, Tuple[float, float]]] = []
        for k, cut in enumerate(cuts):
            down_outer, up_outer = outer[2 * k], outer[2 * k + 1]
            up_outer[0] = down_outer[0]
            up_outer[-1] = down_outer[-1]
            E[k]["down_outer"] = Edge(down_outer, E[k]["down_edge"].clustering.copy())
            E[k]["up_outer"] = Edge(up_outer, E[k]["up_edge"].clustering.copy())
            corner.append(self._corners(cut, E[k], splines[2 * k].total_length + splines[2 * k + 1].total_length))
        # 4: straight edges between known corners
        wave1 = [("o_j_min", "blade_le", "outer_le", o_cluster), ("o_j_max", "blade_te", "outer_te", o_cluster),
                 ("in_j_max", "in_10", "in_11", uniform), ("in_i_min", "in_00", "in_10", uniform), ("in_i_max", "in_01", "in_11", uniform),
                 ("out_j_max", "out_10", "out_11", uniform), ("out_i_min", "out_00", "out_10", uniform), ("out_i_max", "out_01", "out_11", uniform),
                 ("down_i_max", "down_01", "down_11", uniform), ("down_j_min", "in_11", "down_01", uniform), ("down_j_max", "out_10", "down_11", uniform),
                 ("up_i_max_0", "out_11", "up_mid", uniform), ("up_i_max_1", "up_mid", "up_11", uniform), ("up_j_max", "in_10", "up_11", uniform)]
        self._lines(E, corner, wave1, n, dev)
        # 5: combined edges
        order = list(views)
        combined = Edge.combine_batch([[(E[k][e], a, b) for e, a, b in views[name]] for k in range(len(cuts)) for name in order], dev)
        for k in range(len(cuts)):
            for q, name in enumerate(order):
                E[k][name] = combined[k * len(order) + q]
                assert len(E[k][name].points) == n[name]
        # 6: upstream / downstream, from the end points of the combined edges (as computed, O4H.zig:362-420)
        for k, cut in enumerate(cuts):
            cr, le = corner[k], corner[k]["blade_le"]
            inlet, outlet = cr["_inlet"][0], cr["_outlet"][0]
            us, ds = E[k]["upstream_j_max"].points, E[k]["downstream_j_min"].points
            cr["us_10"], cr["us_11"] = (float(us[0, 0]), float(us[0, 1])), (float(us[-1, 0]), float(us[-1, 1]))
            cr["us_00"] = (le[0] - inlet, le[1] - 0.5 * cut.pitch)
            cr["us_01"] = (le[0] - inlet, le[1] + 0.5 * cut.pitch)
            cr["ds_00"], cr["ds_01"] = (float(ds[0, 0]), float(ds[0, 1])), (float(ds[-1, 0]), float(ds[-1, 1]))
            cr["ds_10"] = (cr["ds_00"][0] + outlet, cr["ds_00"][1] + 0.0)
            cr["ds_11"] = (cr["ds_10"][0] + 0.0, cr["ds_10"][1] + cut.pitch)
        wave2 = [("upstream_j_min", "us_00", "us_01", uniform), ("upstream_i_min", "us_00", "us_10", uniform), ("upstream_i_max", "us_01", "us_11", uniform),
                 ("downstream_j_max", "ds_10", "ds_11", uniform), ("downstream_i_min", "ds_00", "ds_10", uniform), ("downstream_i_max", "ds_01", "ds_11", uniform)]
        self._lines(E, corner, wave2, n, dev)
        # the mesh: 8 blocks per cut, connections and conditions with the block indices of the cut
        mesh, groups = Mesh(), []
        for k, cut in enumerate(cuts):
            first = len(mesh.blocks)
            for name in BLOCKS:
                mesh.add_block(f"cut{k}_{name}" if len(cuts) > 1 else name, EdgeBlock(*(E[k][e] for e in _BLOCK_EDGES[name])))
            conns, conds = connections(c, cut.pitch, first)
            mesh.connections.extend(conns)
            mesh.boundary_conditions.extend(conds)
            groups.append((first, first + 1))
        return mesh, groups

    def _corners(self, cut: Cut, e: Dict[str, Edge], profile_length: float) -> Dict[str, Tuple[float, float]]:
        """Corner points known once the blade and its offset exist (O4H.zig:78-346)."""
        c = self.cells
        pt = lambda a: (float(a[0]), float(a[1]))  # noqa: E731
        spacing = profile_length / float(c.up + c.down)
        inlet = self.inlet_distance if self.inlet_distance is not None else spacing * float(c.upstream_i)
        outlet = self.outlet_distance if self.outlet_distance is not None else spacing * float(c.downstream_i)
        le, te, pitch = pt(e["up_edge"].points[0]), pt(e["up_edge"].points[-1]), cut.pitch
        up_o, down_o = e["up_outer"].points, e["down_outer"].points
        x_in, x_out = le[0] - inlet * 0.5, outlet * 0.5 + te[0]
        return {"_inlet": (inlet, 0.0), "_outlet": (outlet, 0.0),
                "blade_le": le, "blade_te": te, "outer_le": pt(up_o[0]), "outer_te": pt(up_o[-1]),
                "in_00": pt(up_o[c.in_up_j]), "in_01": pt(down_o[c.in_down_j]),
                "in_10": (x_in, le[1] + pitch * 0.25), "in_11": (x_in, le[1] - pitch * 0.25),
                "out_00": pt(down_o[c.in_down_j + c.middle_i]), "out_01": pt(up_o[c.in_up_j + c.bulge + c.middle_i + c.out_i]),
                "out_10": (x_out, te[1] - pitch * 0.25), "out_11": (x_out, te[1] + pitch * 0.25),
                "down_01": (le[0] - 0.0, le[1] - 0.5 * pitch), "down_11": (te[0] - 0.0, te[1] - 0.5 * pitch),
                "up_11": (le[0] + 0.0, le[1] + 0.5 * pitch), "up_mid": (te[0] + 0.0, te[1] + 0.5 * pitch)}

    @staticmethod
    def _lines(E, corner, table, n, dev) -> None:
        edges = Edge.init_batch([(n[name], _Line(corner[k][a], corner[k][b]), cl) for k in range(len(E)) for name, a, b, cl in table], dev)
        for k in range(len(E)):
            for q, (name, _, _, _) in enumerate(table):
                E[k][name] = edges[k * len(table) + q]
