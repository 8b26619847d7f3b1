"""Config 4 as BASELINE.json names it: pitch-wise O4H passages around a blade profile (workload generator of bench.py and tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from turbomesh_b200.boundary import Condition, Connection, Range
from turbomesh_b200.clustering import Roberts
from turbomesh_b200.discrete import Edge, Mesh
from turbomesh_b200.synthetic import EdgeBlock

from .geometry import Geometry, Profile
from .templates import O4H, NumCells


def o4h_passages(profile_up: np.ndarray, profile_down: np.ndarray, pitch: float, n_passages: int = 8, factor: int = 48,
                 num_cells=None, o_grid_delta_s: Optional[float] = None, o_grid_thickness: Optional[float] = None) -> Tuple[Mesh, List[int]]:
    """Config 4 as BASELINE.json names it: `n_passages` pitch-wise copies of the 8-block O4H topology (O4H.zig:423-521) around a
    blade profile, passage k shifted by k * pitch; the three pitch-wise periodic connections of the template (#18 upstream,
    #19 down <-> up, #20 downstream; O4H.zig:503-514) are rewired passage k <-> k + 1 as ordinary connections and stay
    periodic (n_passages * pitch) between the last and the first passage.  Cell counts = the T106 example's
    (examples/T106/T106.json) times `factor` (48 -> ~58 M nodes per passage, 463 M nodes at 8 passages).

    `profile_up` / `profile_down`: points on the suction / pressure side from the leading to the trailing edge (e.g. the blade
    edges of the committed T106 fixture).  Returns the mesh (blocks known by their edges only) and the owner list (one
    passage per rank).  The O-grid clustering scales with the refinement ((n - 1) delta_s <= 1, clustering.zig:68-76).
    """

    base = num_cells or NumCells(o_grid=40, middle_i=100, in_up_j=30, in_down_j=10, in_i=10, out_up_j=40, out_down_j=10, out_i=10,
                                 down_j=40, bulge=40, upstream_i=20, downstream_i=10)   # examples/T106/T106.json
    nc = base.scaled(factor)
    geom = Geometry(pitch, Profile(profile_down, profile_up))
    calls = []

    def record(*args):
        calls.append([np.array(a, dtype=np.float64, copy=True) for a in args])
        return np.zeros((len(args[4]), 1, 2))   # placeholder: the blocks are produced on the device from the recorded edges

    tmpl = O4H(blade_clustering=Roberts(0.5, 1.03), num_cells=nc)
    tmpl.o_grid_delta_s = o_grid_delta_s if o_grid_delta_s is not None else 0.5 / nc.o_grid
    if o_grid_thickness is not None:
        tmpl.o_grid_thickness = o_grid_thickness
    one = tmpl.run(geom, tfi=record)
    nb = len(calls)
    assert nb == 8
    BLADE_UP, BLADE_DOWN, IN, OUT, DOWN, UP, UPSTREAM, DOWNSTREAM = range(8)
    mesh = Mesh()
    shift = np.array([0.0, pitch])
    prev = None
    for k in range(n_passages):
        edges = [[a.copy() for a in c] for c in calls]
        for c in edges:
            for a in c[:4]:
                a += k * shift
        if prev is not None:
            # the lower sides of this passage ARE the upper sides of the previous one, bit for bit (connectionDataCheck wants
            # the copies within 1e-15; recomputing x + k * pitch on both sides would differ by an ulp)
            edges[UPSTREAM][0] = prev[UPSTREAM][1].copy()                       # upstream i_min := previous upstream i_max
            edges[DOWNSTREAM][0] = prev[DOWNSTREAM][1].copy()
            n_dn = len(edges[DOWN][1])
            edges[DOWN][1] = prev[UP][1][::-1][:n_dn].copy()                      # down i_max := previous up i_max (its last n_dn points, reversed)
            # corner values of the adjacent j edges follow (tfi.zig:150-162 wants the four corners consistent)
            edges[UPSTREAM][2][0], edges[UPSTREAM][3][0] = edges[UPSTREAM][0][0], edges[UPSTREAM][0][-1]
            edges[DOWNSTREAM][2][0], edges[DOWNSTREAM][3][0] = edges[DOWNSTREAM][0][0], edges[DOWNSTREAM][0][-1]
            edges[DOWN][2][-1], edges[DOWN][3][-1] = edges[DOWN][1][0], edges[DOWN][1][-1]
        for b, c in enumerate(edges):
            mesh.add_block(f"p{k}_{one.names[b]}", EdgeBlock(Edge(c[0], c[4]), Edge(c[1], c[5]), Edge(c[2], c[6]), Edge(c[3], c[7])))
        prev = edges
    periodic = [c for c in one.connections if c.periodicity is not None]
    assert len(periodic) == 3
    for k in range(n_passages):
        off = k * nb
        for c in one.connections:
            if c.periodicity is not None:
                continue
            r0, r1 = c.ranges
            mesh.connections.append(Connection((Range(r0.block + off, r0.side, r0.start, r0.end), Range(r1.block + off, r1.side, r1.start, r1.end))))
        for bc in one.boundary_conditions:
            r = bc.range
            mesh.boundary_conditions.append(Condition(Range(r.block + off, r.side, r.start, r.end), bc.kind))
    for c in periodic:   # template: ranges[0] = lower side (i_min / down:i_max), ranges[1] = upper side, x0 + (0, pitch) == x1
        lo, hi = c.ranges
        for k in range(n_passages - 1):   # upper side of passage k == lower side of passage k + 1
            mesh.connections.append(Connection((Range(hi.block + k * nb, hi.side, hi.start, hi.end), Range(lo.block + (k + 1) * nb, lo.side, lo.start, lo.end))))
        mesh.connections.append(Connection((Range(lo.block, lo.side, lo.start, lo.end), Range(hi.block + (n_passages - 1) * nb, hi.side, hi.start, hi.end)),
                                           (0.0, n_passages * pitch)))
    owner = [k for k in range(n_passages) for _ in range(nb)]
    return mesh, owner
