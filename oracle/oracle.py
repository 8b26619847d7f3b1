"""ctypes front-end of the CPU oracle (``oracle/turbomesh_oracle.c``).  TEST INFRASTRUCTURE ONLY.

The oracle is the checker for the CUDA path and the timed CPU baseline of ``bench.py``; nothing under
``turbomesh_b200/`` may import it.  PARITY UNPINNED: see the C file's header.

Meshes are duck-typed: ``mesh.blocks[k].points`` float64 arrays (ni, nj, 2); ``mesh.connections[k]`` with
``.ranges[2]`` (``.block .side .start .end``) and ``.periodicity``; ``mesh.boundary_conditions[k]`` with
``.range`` and ``.kind`` -- i.e. the reference's ``discrete.Mesh`` (``discrete.zig:166-195``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libturbomesh_oracle.so")

SOLVER_GMRES, SOLVER_BICGSTAB = 0, 1
PRECOND_DIAGONAL, PRECOND_ILU0 = 0, 1
CF_LAPLACE, CF_WHITE = 0, 1
KIND_NAMES = ("fixed", "smoothed", "connected", "laplacian_smoothed", "sliding_circ")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "turbomesh_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libturbomesh_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class Block(C.Structure):
    _fields_ = [("ni", C.c_uint64), ("nj", C.c_uint64), ("xy", C.POINTER(C.c_double))]


class Range(C.Structure):
    _fields_ = [("block", C.c_uint64), ("side", C.c_uint32), ("_pad", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64)]


class Connection(C.Structure):
    _fields_ = [("ranges", Range * 2), ("has_periodicity", C.c_int32), ("_pad", C.c_int32), ("periodicity", C.c_double * 2)]


class Condition(C.Structure):
    _fields_ = [("range", Range), ("kind", C.c_uint32), ("_pad", C.c_uint32)]


class Options(C.Structure):
    _fields_ = [("solver", C.c_int32), ("preconditioner", C.c_int32), ("control_function", C.c_int32), ("restart", C.c_int32),
                ("max_iters", C.c_uint64), ("rtol", C.c_double), ("atol", C.c_double), ("ds_target", C.c_double), ("theta_target", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("outer_iterations", C.c_uint64), ("krylov_iterations", C.c_uint64), ("matvecs", C.c_uint64), ("precond_applies", C.c_uint64),
                ("not_converged", C.c_uint64), ("last_sumsq_x", C.c_double), ("last_sumsq_y", C.c_double), ("last_residual", C.c_double),
                ("last_max_update", C.c_double), ("seconds_fill", C.c_double), ("seconds_solve", C.c_double), ("seconds_total", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.orc_last_error.restype = C.c_char_p
        L.orc_tfi.argtypes = [C.c_uint64, C.c_uint64] + [dp] * 9
        L.orc_options_default.argtypes = [C.POINTER(Options)]
        L.orc_system_create.argtypes = [C.POINTER(Block), C.c_size_t, C.POINTER(Connection), C.c_size_t, C.POINTER(Condition), C.c_size_t,
                                        C.POINTER(Options), C.POINTER(C.c_void_p)]
        L.orc_system_destroy.argtypes = [C.c_void_p]
        L.orc_system_fill.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_system_fill_specific.argtypes = [C.c_void_p, C.c_int]
        L.orc_system_iterate.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_smooth_mesh.argtypes = [C.POINTER(Block), C.c_size_t, C.POINTER(Connection), C.c_size_t, C.POINTER(Condition), C.c_size_t,
                                      C.c_uint64, C.POINTER(Options), C.POINTER(Stats)]
        for name in ("dof", "nnz", "n_boundary", "n_junctions"):
            f = getattr(L, "orc_system_" + name); f.argtypes = [C.c_void_p]; f.restype = C.c_uint64
        for name, rt in (("lhs_p", ip), ("lhs_i", ip), ("lhs_values", dp), ("rhs_x", dp), ("rhs_y", dp), ("control_function", dp), ("kinds", C.POINTER(C.c_uint8))):
            f = getattr(L, "orc_system_" + name); f.argtypes = [C.c_void_p]; f.restype = rt
        L.orc_system_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.orc_system_junction.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), dp, ip, dp]
        L.orc_csr_solve.argtypes = [C.c_uint64, ip, ip, dp, dp, dp, C.POINTER(Options), C.POINTER(Stats)]
        L.orc_block_to_soa.argtypes = [C.c_uint64, C.c_uint64, dp, dp, dp]
        L.orc_block_to_soa.restype = None
        L.orc_viewer_points.argtypes = [C.POINTER(Block), C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_viewer_points.restype = None
        L.orc_viewer_index_count.argtypes = [C.POINTER(Block), C.c_size_t]
        L.orc_viewer_index_count.restype = C.c_uint64
        L.orc_viewer_wireframe.argtypes = [C.POINTER(Block), C.c_size_t, C.POINTER(C.c_uint32)]
        L.orc_viewer_wireframe.restype = None
        L.orc_clustering.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64, dp]
        L.orc_line_interpolate.argtypes = [dp, dp, dp, C.c_uint64, dp]
        L.orc_line_interpolate.restype = None
        L.orc_spline_interpolate.argtypes = [C.c_uint64, dp, dp, dp, dp, C.c_uint64, dp, C.c_double, dp, C.c_uint64, dp]
        L.orc_spline_interpolate.restype = None
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(f"oracle error {rc}: {lib().orc_last_error().decode()}")


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def options(solver="gmres", preconditioner="ilu0", control_function="laplace", rtol=1e-6, atol=1e-8, max_iters=1000, restart=30,
            ds_target=1e-6, theta_target=0.5 * np.pi) -> Options:
    """Defaults = the reference's (GMRES.zig:21-24, examples/T106/T106.json:29-33)."""
    o = Options()
    lib().orc_options_default(C.byref(o))
    o.solver = {"gmres": SOLVER_GMRES, "bicgstab": SOLVER_BICGSTAB}[solver]
    o.preconditioner = {"diagonal": PRECOND_DIAGONAL, "ilu0": PRECOND_ILU0}[preconditioner]
    o.control_function = {"laplace": CF_LAPLACE, "white": CF_WHITE}[control_function]
    o.rtol, o.atol, o.max_iters, o.restart = rtol, atol, max_iters, restart
    o.ds_target, o.theta_target = ds_target, theta_target
    return o


def tight_options(**kw) -> Options:
    """'Exact Picard step' mode used for parity: the same solver, tolerances tightened."""
    kw.setdefault("rtol", 1e-15)
    kw.setdefault("atol", 1e-15)
    kw.setdefault("max_iters", 200000)
    return options(**kw)


def tfi(x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2) -> np.ndarray:
    """``tfi.linear2dBoundaryBlendedControlFunction`` (tfi.zig:112-208) on the CPU; returns (ni, nj, 2)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2)]
    n, m = len(arrs[4]), len(arrs[6])
    assert arrs[0].shape == (n, 2) and arrs[1].shape == (n, 2) and arrs[2].shape == (m, 2) and arrs[3].shape == (m, 2)
    assert len(arrs[5]) == n and len(arrs[7]) == m
    out = np.empty((n, m, 2), dtype=np.float64)
    _check(lib().orc_tfi(n, m, *[_dp(a) for a in arrs], _dp(out)))
    return out


@dataclass
class _CMesh:
    blocks: object
    conns: object
    bcs: object
    arrays: list
    nb: int
    nc: int
    nbc: int


def _to_c(mesh) -> _CMesh:
    nb, nc, nbc = len(mesh.blocks), len(mesh.connections), len(mesh.boundary_conditions)
    arrays = []
    cb = (Block * max(nb, 1))()
    for k, b in enumerate(mesh.blocks):
        a = b.points
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.ndim == 3 and a.shape[2] == 2
        arrays.append(a)
        cb[k].ni, cb[k].nj, cb[k].xy = a.shape[0], a.shape[1], _dp(a)
    cc = (Connection * max(nc, 1))()
    for k, c in enumerate(mesh.connections):
        for s in range(2):
            r = c.ranges[s]
            cc[k].ranges[s].block, cc[k].ranges[s].side, cc[k].ranges[s].start, cc[k].ranges[s].end = r.block, int(r.side), r.start, r.end
        if c.periodicity is not None:
            cc[k].has_periodicity = 1
            cc[k].periodicity[0], cc[k].periodicity[1] = c.periodicity
    cd = (Condition * max(nbc, 1))()
    for k, bc in enumerate(mesh.boundary_conditions):
        r = bc.range
        cd[k].range.block, cd[k].range.side, cd[k].range.start, cd[k].range.end = r.block, int(r.side), r.start, r.end
        cd[k].kind = int(bc.kind)
    return _CMesh(cb, cc, cd, arrays, nb, nc, nbc)


def smooth_mesh(mesh, iterations: int, opts: Options = None) -> dict:
    """``smoothing.smooth.mesh`` (smooth.zig:74-166) on the CPU; smooths ``mesh`` in place, returns the stats."""
    opts = opts or options()
    cm = _to_c(mesh)
    st = Stats()
    _check(lib().orc_smooth_mesh(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, iterations, C.byref(opts), C.byref(st)))
    return st.as_dict()


class System:
    """Step-by-step access to the restated ``RowCompressedMatrixSystem2d`` (smooth.zig:277-1166)."""

    def __init__(self, mesh, opts: Options = None):
        self.opts = opts or options()
        self._cm = _to_c(mesh)
        self.mesh = mesh
        h = C.c_void_p()
        cm = self._cm
        _check(lib().orc_system_create(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, C.byref(self.opts), C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib().orc_system_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dof(self):
        return int(lib().orc_system_dof(self._h))

    @property
    def nnz(self):
        return int(lib().orc_system_nnz(self._h))

    def fill(self, iteration: int = 0):
        _check(lib().orc_system_fill(self._h, iteration))

    def fill_specific(self, y_mode: bool):
        lib().orc_system_fill_specific(self._h, 1 if y_mode else 0)

    def iterate(self, n: int):
        _check(lib().orc_system_iterate(self._h, n))

    def stats(self) -> dict:
        st = Stats()
        lib().orc_system_stats(self._h, C.byref(st))
        return st.as_dict()

    def csr(self):
        """(indptr, indices, values, rhs_x, rhs_y) copies of the assembled system."""
        L, dof, nnz = lib(), self.dof, self.nnz
        p = np.ctypeslib.as_array(L.orc_system_lhs_p(self._h), (dof + 1,)).copy()
        i = np.ctypeslib.as_array(L.orc_system_lhs_i(self._h), (nnz,)).copy()
        v = np.ctypeslib.as_array(L.orc_system_lhs_values(self._h), (nnz,)).copy()
        rx = np.ctypeslib.as_array(L.orc_system_rhs_x(self._h), (dof,)).copy()
        ry = np.ctypeslib.as_array(L.orc_system_rhs_y(self._h), (dof,)).copy()
        return p, i, v, rx, ry

    def control_function(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().orc_system_control_function(self._h), (self.dof, 2)).copy()

    def kinds(self) -> np.ndarray:
        """Node kinds in the reference's flat boundary numbering (boundary.zig:248-285)."""
        n = int(lib().orc_system_n_boundary(self._h))
        return np.ctypeslib.as_array(lib().orc_system_kinds(self._h), (n,)).copy()

    def junctions(self):
        out = []
        for l in range(int(lib().orc_system_n_junctions(self._h))):
            ids = (C.c_uint64 * 4)(); per = (C.c_double * 8)(); st = (C.c_int32 * 6)(); rhs = (C.c_double * 2)()
            lib().orc_system_junction(self._h, l, ids, per, st, rhs)
            k = [int(v) for v in ids if v != 2**64 - 1]
            out.append({"ids": k, "periodicity": [(per[2 * q], per[2 * q + 1]) for q in range(len(k))],
                        "stencil": [int(v) for v in st if v >= 0], "rhs": (rhs[0], rhs[1])})
        return out


def csr_solve(indptr, indices, values, rhs, x0=None, opts: Options = None):
    """Runs the restated GMRES / BiCGStab on an arbitrary CSR system (known-answer tests)."""
    opts = opts or options()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    values = np.ascontiguousarray(values, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.zeros_like(rhs) if x0 is None else np.array(x0, dtype=np.float64)
    st = Stats()
    ip = C.POINTER(C.c_int32)
    _check(lib().orc_csr_solve(len(rhs), indptr.ctypes.data_as(ip), indices.ctypes.data_as(ip), _dp(values), _dp(rhs), _dp(x), C.byref(opts), C.byref(st)))
    return x, st.as_dict()


def block_to_soa(points: np.ndarray):
    """CoordinateX / CoordinateY arrays of a block as the reference's CGNS writer lays them out (cgns.zig:69-101): i fastest."""
    a = np.ascontiguousarray(points, dtype=np.float64)
    ni, nj = a.shape[0], a.shape[1]
    x, y = np.empty(ni * nj), np.empty(ni * nj)
    lib().orc_block_to_soa(ni, nj, _dp(a), _dp(x), _dp(y))
    return x, y


def clustering(kind: str, n: int, alpha: float = 0.0, beta: float = 0.0, delta_s: float = 0.0) -> np.ndarray:
    """``clustering.create`` (clustering.zig:9-116): kind = uniform | roberts | single_hyperbolic_clustering."""
    out = np.empty(n)
    _check(lib().orc_clustering({"uniform": 0, "roberts": 1, "single_hyperbolic_clustering": 2}[kind], alpha, beta, delta_s, n, _dp(out)))
    return out


def line_interpolate(start, end, u) -> np.ndarray:
    """``Line.interpolate`` (geometry.zig:26-40)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    a, b = np.array(start, dtype=np.float64), np.array(end, dtype=np.float64)
    out = np.empty((len(u), 2))
    lib().orc_line_interpolate(_dp(a), _dp(b), _dp(u), len(u), _dp(out))
    return out


def spline_interpolate(params, points, zx, zy, sample_arc, total_length: float, u) -> np.ndarray:
    """``FittingSpline.interpolate`` (spline.zig:74-81, 112-139, 202-222) on the tables of an already fitted spline."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (params, points, zx, zy, sample_arc, u)]
    out = np.empty((len(arrs[5]), 2))
    lib().orc_spline_interpolate(len(arrs[0]), _dp(arrs[0]), _dp(arrs[1]), _dp(arrs[2]), _dp(arrs[3]), len(arrs[4]), _dp(arrs[4]), float(total_length),
                                 _dp(arrs[5]), len(arrs[5]), _dp(out))
    return out


def viewer_buffers(blocks_points):
    """``createPointBuffer`` + ``createWireframeElementBuffer`` (src/gui/lib.zig:227-318) for a list of (ni, nj, 2) blocks:
    returns (points f32 [2N], range_x f32[2], range_y f32[2], indices u32)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in blocks_points]
    blocks = (Block * len(arrs))()
    for k, a in enumerate(arrs):
        blocks[k].ni, blocks[k].nj, blocks[k].xy = a.shape[0], a.shape[1], _dp(a)
    n = sum(a.shape[0] * a.shape[1] for a in arrs)
    pts = np.empty(2 * n, dtype=np.float32)
    rx, ry = np.empty(2, dtype=np.float32), np.empty(2, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_viewer_points(blocks, len(arrs), pts.ctypes.data_as(fp), rx.ctypes.data_as(fp), ry.ctypes.data_as(fp))
    idx = np.empty(int(lib().orc_viewer_index_count(blocks, len(arrs))), dtype=np.uint32)
    lib().orc_viewer_wireframe(blocks, len(arrs), idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return pts, rx, ry, idx
