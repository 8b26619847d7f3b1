"""Discrete mesh model -- host-side mirror of ``src/core/discrete.zig``.

``Edge`` / ``Block2d`` / ``Mesh`` keep the reference's field names.  ``Block2d.init`` is one of the two call sites of the accelerated path: it allocates the block and runs
the boundary-blended TFI (``discrete.zig:142-159`` -> ``tfi.zig:112-208``) -- here through the C ABI
(``tm_tfi_block``) on the GPU.  There is no CPU fallback; a ``tfi=`` callable can be injected (the test
suite injects the CPU oracle as the checker).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import clustering as cluster
from .boundary import Condition, Connection


def _addr(a: np.ndarray) -> int:
    """address of the first element of a contiguous array (cheaper than ``a.ctypes`` when there are thousands of them)"""
    return a.__array_interface__["data"][0]


class Edge:
    """``discrete.zig:12-91``: points (n,2) + clustering (n,) of one block edge."""

    def __init__(self, points: np.ndarray, clustering: np.ndarray):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.clustering = np.ascontiguousarray(clustering, dtype=np.float64)
        assert self.points.shape == (len(self.clustering), 2)

    @staticmethod
    def init(n: int, curve, clustering) -> "Edge":
        """``Edge.init``, ``discrete.zig:17-31``."""
        u = cluster.create(clustering, n)
        if not hasattr(curve, "interpolate"):
            raise TypeError("curve must offer interpolate(u) (geometry.zig:18-40 Line, spline.zig:24-222 FittingSpline)")
        data = curve.interpolate(u)
        return Edge(data, u)

    @staticmethod
    def init_batch(specs, device: int = -1) -> List["Edge"]:
        """``Edge.init`` for many edges at once on the GPU (``tm_edges_discretize``): ``specs`` = [(n, curve, clustering)].

        One launch for all edges (the batch-of-cuts pipeline discretises ~30 edges per cut); the curve arithmetic is
        bit-exact with :meth:`init`, the Roberts / hyperbolic clusterings agree to a few ulp (CUDA's pow / tanh)."""
        import ctypes as C

        from . import _lib

        jobs = (_lib.TmEdgeJob * max(len(specs), 1))()
        total = sum(int(n) for n, _, _ in specs)
        pts_all, cl_all = np.empty((total, 2), dtype=np.float64), np.empty(total, dtype=np.float64)   # one buffer for the batch, an edge = a slice
        p0, c0 = _addr(pts_all), _addr(cl_all)
        keep, splines, off = [], {}, 0
        bounds = []
        for k, (n, curve, clustering) in enumerate(specs):
            j = jobs[k]
            n = int(n)
            j.n = n
            j.points, j.clustering = p0 + 16 * off, c0 + 8 * off
            bounds.append((off, off + n))
            off += n
            if isinstance(clustering, cluster.Uniform):
                j.clustering_kind = 0
            elif isinstance(clustering, cluster.Roberts):
                j.clustering_kind, j.alpha, j.beta = 1, clustering.alpha, clustering.beta
            elif isinstance(clustering, cluster.SingleHyperbolicClustering):
                j.clustering_kind, j.delta_s = 2, clustering.delta_s
            else:
                raise TypeError("unknown clustering function")
            if hasattr(curve, "start") and hasattr(curve, "end"):       # a line (geometry.zig:18-40)
                j.curve_kind = 0
                j.line_start[0], j.line_start[1], j.line_end[0], j.line_end[1] = curve.start[0], curve.start[1], curve.end[0], curve.end[1]
            elif hasattr(curve, "second_derivs"):                        # an already fitted spline (spline.zig:24-40): its tables are borrowed
                j.curve_kind = 1
                if id(curve) not in splines:
                    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (curve.params, curve.points, curve.second_derivs[0], curve.second_derivs[1], curve.sample_arc)]
                    sp = _lib.TmSpline()
                    sp.n_points, sp.n_samples, sp.total_length = len(arrs[0]), len(arrs[4]), float(curve.total_length)
                    sp.params, sp.points, sp.second_derivs_x, sp.second_derivs_y, sp.sample_arc = [a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs]
                    keep.append(arrs)
                    splines[id(curve)] = sp
                j.spline = C.pointer(splines[id(curve)])
            else:
                raise TypeError("curve must be a line (start, end) or a fitted spline (params, points, second_derivs, sample_arc, total_length)")
        _lib.check(_lib.load().tm_edges_discretize(jobs, len(specs), device))
        return [Edge(pts_all[a:b], cl_all[a:b]) for a, b in bounds]

    def copy(self) -> "Edge":
        return Edge(self.points.copy(), self.clustering.copy())

    @staticmethod
    def combine_batch(jobs, device: int = -1) -> List["Edge"]:
        """``Edge.combine`` (``discrete.zig:38-91``) for many edges at once on the GPU (``tm_edges_combine``): ``jobs`` = one list
        of views ``(edge, start, end)`` per combined edge (``EdgeView``, ``discrete.zig:94-136``)."""
        import ctypes as C

        from . import _lib

        cj = (_lib.TmCombineJob * max(len(jobs), 1))()
        n_views = sum(len(views) for views in jobs)
        cv = (_lib.TmEdgeView * max(n_views, 1))()            # the views of all jobs in one array
        v0, vsize = C.addressof(cv), C.sizeof(_lib.TmEdgeView)
        sizes = [sum(abs(int(a) - int(b)) + 1 for _, a, b in views) - (len(views) - 1) for views in jobs]
        total = sum(sizes)
        pts_all, cl_all = np.empty((total, 2), dtype=np.float64), np.empty(total, dtype=np.float64)
        p0, c0 = _addr(pts_all), _addr(cl_all)
        off = q = 0
        bounds = []
        for k, views in enumerate(jobs):
            cj[k].views, cj[k].n_views, cj[k].points, cj[k].clustering = v0 + q * vsize, len(views), p0 + 16 * off, c0 + 8 * off
            for edge, start, end in views:
                v = cv[q]
                v.points, v.clustering, v.n, v.start, v.end = _addr(edge.points), _addr(edge.clustering), len(edge.clustering), int(start), int(end)
                q += 1
            bounds.append((off, off + sizes[k]))
            off += sizes[k]
        _lib.check(_lib.load().tm_edges_combine(cj, len(jobs), device))
        return [Edge(pts_all[a:b], cl_all[a:b]) for a, b in bounds]

    @staticmethod
    def project_normal_batch(jobs, device: int = -1) -> List[np.ndarray]:
        """``projectNormal`` (``templates/O4H.zig:531-574``) for many edges at once on the GPU: ``jobs`` = [(points (n, 2), distance)]."""
        from . import _lib

        cj = (_lib.TmProjectJob * max(len(jobs), 1))()
        srcs = [np.ascontiguousarray(points, dtype=np.float64) for points, _ in jobs]
        out_all = np.empty((sum(len(a) for a in srcs), 2), dtype=np.float64)
        o0, off, bounds = _addr(out_all), 0, []
        for k, (src, (_, distance)) in enumerate(zip(srcs, jobs)):
            cj[k].points, cj[k].n, cj[k].distance, cj[k].out = _addr(src), len(src), float(distance), o0 + 16 * off
            bounds.append((off, off + len(src)))
            off += len(src)
        _lib.check(_lib.load().tm_edges_project_normal(cj, len(jobs), device))
        return [out_all[a:b] for a, b in bounds]


class FittedSpline:
    """Tables of a fitted ``spline.FittingSpline(2)`` (``spline.zig:24-40``) produced on the GPU by ``tm_splines_fit``; usable as the
    curve of :meth:`Edge.init_batch`."""

    def __init__(self, points, params, zx, zy, sample_arc, total_length):
        self.points, self.params, self.second_derivs, self.sample_arc, self.total_length = points, params, [zx, zy], sample_arc, total_length

    @staticmethod
    def fit_batch(point_sets, n_samples: int = 201, device: int = -1) -> List["FittedSpline"]:
        import ctypes as C

        from . import _lib

        dp = C.POINTER(C.c_double)
        jobs = (_lib.TmSplineFitJob * max(len(point_sets), 1))()
        srcs = [np.ascontiguousarray(pts, dtype=np.float64) for pts in point_sets]
        total = sum(len(a) for a in srcs)
        # one buffer per table for the whole batch (the library then brings each back in one copy); a spline = slices of them
        params, zx, zy = np.empty(total), np.empty(total), np.empty(total)
        arcs, lengths = np.empty(len(srcs) * n_samples), np.empty(max(len(srcs), 1))
        off, bounds = 0, []
        for k, src in enumerate(srcs):
            n = len(src)
            j = jobs[k]
            j.n_points, j.points, j.n_samples = n, src.ctypes.data_as(dp), n_samples
            j.params, j.second_derivs_x, j.second_derivs_y = [C.cast(_addr(a) + 8 * off, dp) for a in (params, zx, zy)]
            j.sample_arc, j.total_length = C.cast(_addr(arcs) + 8 * k * n_samples, dp), C.cast(_addr(lengths) + 8 * k, dp)
            bounds.append((off, off + n))
            off += n
        _lib.check(_lib.load().tm_splines_fit(jobs, len(srcs), device))
        return [FittedSpline(src, params[a:b], zx[a:b], zy[a:b], arcs[k * n_samples:(k + 1) * n_samples], float(lengths[k])) for k, (src, (a, b)) in enumerate(zip(srcs, bounds))]


TfiFn = Callable[..., np.ndarray]


def _default_tfi(*args) -> np.ndarray:
    from .smoothing import tfi_block  # GPU back-end (C ABI); raises if the CUDA library is unusable

    return tfi_block(*args)


class Block2d:
    """``discrete.zig:138-164``; ``points`` is the ``Mat2d`` view: array (ni, nj, 2), j fastest."""

    def __init__(self, points: np.ndarray):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        assert self.points.ndim == 3 and self.points.shape[2] == 2

    @property
    def size(self):
        return self.points.shape[0], self.points.shape[1]

    @staticmethod
    def init(i_min: Edge, i_max: Edge, j_min: Edge, j_max: Edge, tfi: Optional[TfiFn] = None) -> "Block2d":
        assert len(i_min.points) == len(i_max.points)
        assert len(j_min.points) == len(j_max.points)
        fn = tfi or _default_tfi
        pts = fn(i_min.points, i_max.points, j_min.points, j_max.points,
                 i_min.clustering, i_max.clustering, j_min.clustering, j_max.clustering)
        return Block2d(pts)


@dataclass
class Mesh:
    """``discrete.zig:166-195``."""

    blocks: List[Block2d] = field(default_factory=list)
    names: List[str] = field(default_factory=list)
    connections: List[Connection] = field(default_factory=list)
    boundary_conditions: List[Condition] = field(default_factory=list)

    def add_block(self, name: str, block: Block2d) -> int:
        self.blocks.append(block)
        self.names.append(name)
        return len(self.blocks) - 1

    def num_nodes(self) -> int:
        return sum(b.size[0] * b.size[1] for b in self.blocks)

    def copy(self) -> "Mesh":
        return Mesh([Block2d(b.points.copy()) for b in self.blocks], list(self.names), list(self.connections), list(self.boundary_conditions))
