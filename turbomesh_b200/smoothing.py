"""Smoothing front-end -- host-side mirror of ``src/core/smoothing/`` routed through the C ABI.

* ``tfi_block``      -> ``tm_tfi_block``    (``tfi.linear2dBoundaryBlendedControlFunction``, tfi.zig:112-208)
* ``mesh`` / ``smooth_mesh`` -> ``tm_smooth_mesh`` (``smoothing.smooth.mesh``, smooth.zig:74-166)
* ``DeviceMesh``     -> the ``tm_mesh_*`` handle API (device-resident meshes for benchmarks and batches)

Everything here needs the CUDA library and a GPU; nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import TmBlock, TmCondition, TmConnection, TmSmoothOptions, TmSmoothStats, check


# ---- solver.Option / wall_control_function.Algorithm mirrors -------------------------------------
@dataclass
class White:
    """``wall_control_function.White`` (wall_control_function.zig:56-61)."""

    ds_target: float
    theta_target: float = 0.5 * math.pi


@dataclass
class Laplace:
    pass


@dataclass
class CudaSolver:
    """The ``"cuda"`` variant added to ``solver.Option`` (solver.zig:18-27); see INTEGRATION.md.

    ``method``: ``"picard_bicgstab"`` reproduces the reference's outer/inner structure (lagged coefficients, linear
    solve per outer iteration, tolerances with the reference's defaults); ``"relax"`` runs ``sweeps_per_iteration``
    damped-Jacobi sweeps of the nonlinear system per outer iteration; ``"multigrid"`` runs V(nu,nu) cycles of a
    geometric FAS multigrid (nu = ``sweeps_per_iteration``) on a single block with fixed boundary nodes.
    """

    method: str = "picard_bicgstab"
    rtol: float = 1e-6
    atol: float = 1e-8
    max_inner_iterations: int = 1000
    omega: float = 1.0
    sweeps_per_iteration: int = 1
    stop_max_update: float = 0.0
    inner_refinement_cycles: int = 0
    fail_on_no_convergence: bool = False
    device: int = -1

    @staticmethod
    def tight(**kw) -> "CudaSolver":
        """'Exact Picard step' settings used for parity against the tight-tolerance oracle.

        The inner tolerance applies to the 2-norm of the row-scaled residual (a length: the Jacobi update), so an
        absolute 1e-13 is ~3 decades above what fp64 can resolve on unit-chord meshes of up to ~1e5 nodes.
        """
        kw.setdefault("rtol", 0.0)
        kw.setdefault("atol", 1e-13)
        kw.setdefault("max_inner_iterations", 200000)
        kw.setdefault("inner_refinement_cycles", 2)
        return CudaSolver(**kw)


def make_options(iterations: int, solver: Optional[CudaSolver] = None, control_function=None) -> TmSmoothOptions:
    solver = solver or CudaSolver()
    control_function = control_function or Laplace()
    o = TmSmoothOptions()
    _lib.load().tm_smooth_options_default(C.byref(o))
    o.solver = {"picard_bicgstab": _lib.TM_SOLVER_PICARD_BICGSTAB, "relax": _lib.TM_SOLVER_RELAX,
                "multigrid": _lib.TM_SOLVER_FAS_MULTIGRID}[solver.method]
    o.iterations = int(iterations)
    o.rtol, o.atol, o.max_inner_iterations = solver.rtol, solver.atol, int(solver.max_inner_iterations)
    o.omega, o.sweeps_per_iteration, o.stop_max_update = solver.omega, int(solver.sweeps_per_iteration), solver.stop_max_update
    o.inner_refinement_cycles = int(solver.inner_refinement_cycles)
    o.fail_on_no_convergence = 1 if solver.fail_on_no_convergence else 0
    o.device = solver.device
    if isinstance(control_function, White):
        o.control_function = _lib.TM_CF_WHITE
        o.white_ds_target, o.white_theta_target = control_function.ds_target, control_function.theta_target
    else:
        o.control_function = _lib.TM_CF_LAPLACE
    return o


def control_function_from_json(obj: dict):
    (tag, val), = obj.items()
    if tag == "laplace":
        return Laplace()
    if tag == "white":
        return White(float(val["ds_target"]), float(val.get("theta_target", 0.5 * math.pi)))
    raise ValueError(f"unknown wall control function {tag!r}")


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _edge_args(x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2)]
    ni, nj = len(arrs[4]), len(arrs[6])
    if arrs[0].shape != (ni, 2) or arrs[1].shape != (ni, 2) or arrs[2].shape != (nj, 2) or arrs[3].shape != (nj, 2) or len(arrs[5]) != ni or len(arrs[7]) != nj:
        raise ValueError("inconsistent edge sizes")
    return arrs, ni, nj


def tfi_block(x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2, out: Optional[np.ndarray] = None) -> np.ndarray:
    """GPU TFI of one block with host buffers (``tm_tfi_block``); returns the ``Mat2d`` view (ni, nj, 2)."""
    arrs, ni, nj = _edge_args(x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2)
    if out is None:
        out = np.empty((ni, nj, 2), dtype=np.float64)
    assert out.shape == (ni, nj, 2) and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
    check(_lib.load().tm_tfi_block(ni, nj, *[_dp(a) for a in arrs], _dp(out)))
    return out


class _CMesh:
    """Flattens a ``discrete.Mesh`` into the C-ABI arrays (keeps the numpy buffers alive)."""

    def __init__(self, mesh, with_coords: bool = True):
        self.nb, self.nc, self.nbc = len(mesh.blocks), len(mesh.connections), len(mesh.boundary_conditions)
        self.arrays = []
        self.blocks = (TmBlock * max(self.nb, 1))()
        for k, b in enumerate(mesh.blocks):
            a = b.points
            if a is None:  # a block known by its edges only (synthetic.EdgeBlock): coordinates come from the device TFI
                if with_coords:
                    raise ValueError(f"block {k} has no coordinates to upload")
                self.blocks[k].ni, self.blocks[k].nj = b.size
                self.blocks[k].xy = None
                continue
            if not (a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.ndim == 3 and a.shape[2] == 2):
                raise ValueError("block points must be C-contiguous float64 arrays of shape (ni, nj, 2)")
            self.arrays.append(a)
            self.blocks[k].ni, self.blocks[k].nj = a.shape[0], a.shape[1]
            self.blocks[k].xy = _dp(a) if with_coords else None
        self.conns = (TmConnection * max(self.nc, 1))()
        for k, c in enumerate(mesh.connections):
            for s in range(2):
                r = c.ranges[s]
                cr = self.conns[k].ranges[s]
                cr.block, cr.side, cr.start, cr.end = r.block, int(r.side), r.start, r.end
            if c.periodicity is not None:
                self.conns[k].has_periodicity = 1
                self.conns[k].periodicity[0], self.conns[k].periodicity[1] = c.periodicity
        self.bcs = (TmCondition * max(self.nbc, 1))()
        for k, bc in enumerate(mesh.boundary_conditions):
            r = bc.range
            cr = self.bcs[k].range
            cr.block, cr.side, cr.start, cr.end = r.block, int(r.side), r.start, r.end
            self.bcs[k].kind = int(bc.kind)


def smooth_mesh(mesh, iterations: int, solver: Optional[CudaSolver] = None, control_function=None) -> dict:
    """``smoothing.smooth.mesh`` (smooth.zig:74-166) on the GPU with host buffers: smooths ``mesh`` in place."""
    opts = make_options(iterations, solver, control_function)
    cm = _CMesh(mesh)
    st = TmSmoothStats()
    check(_lib.load().tm_smooth_mesh(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, C.byref(opts), C.byref(st)))
    return st.as_dict()


mesh = smooth_mesh  # the reference's name: smoothing.smooth.mesh


class DeviceMesh:
    """Device-resident mesh (``tm_mesh_*``): upload / TFI once, smooth repeatedly, download when needed."""

    def __init__(self, mesh, device: int = -1, stream: int = 0, upload: bool = True, owner=None, rank: Optional[int] = None,
                 n_ranks: int = 1, unique_id: Optional[bytes] = None):
        """``owner`` (one rank per block) makes the mesh distributed: ``rank`` = this process' rank (one process per GPU,
        ``unique_id`` from :func:`dist_unique_id` shared by all ranks), or ``rank=None`` to emulate all ``n_ranks`` ranks
        inside this process on one GPU (tests)."""
        self._L = _lib.load()
        self._cm = _CMesh(mesh, with_coords=upload)
        self.mesh = mesh
        h = C.c_void_p()
        cm = self._cm
        sp = C.c_void_p(stream) if stream else None
        if owner is None:
            self.owner, self.rank, self.n_ranks = [0] * cm.nb, 0, 1
            check(self._L.tm_mesh_create(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, device, sp, C.byref(h)))
        else:
            assert len(owner) == cm.nb
            self.owner, self.rank, self.n_ranks = [int(o) for o in owner], rank, int(n_ranks)
            own = (C.c_int32 * cm.nb)(*self.owner)
            uid = (C.c_uint8 * _lib.TM_UNIQUE_ID_BYTES)(*unique_id) if unique_id is not None else None
            check(self._L.tm_mesh_create_distributed(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, own, -1 if rank is None else int(rank),
                                                     int(n_ranks), uid, device, sp, C.byref(h)))
        self._h = h
        self._opts = None

    def owns(self, block: int) -> bool:
        return self.rank is None or self.owner[block] == self.rank

    @property
    def halo_path(self) -> str:
        """How the per-sweep halo exchange travels (``tm_mesh_halo_path``)."""
        return {0: "none", 1: "emulated (one process)", 2: "nccl send/recv", 3: "nvlink peer memory (cuda ipc push)"}[int(self._L.tm_mesh_halo_path(self._h))]

    @property
    def local_node_count(self) -> int:
        return int(self._L.tm_mesh_local_node_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.tm_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def node_count(self) -> int:
        return int(self._L.tm_mesh_node_count(self._h))

    def tfi_block(self, block: int, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2):
        arrs, ni, nj = _edge_args(x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2)
        check(self._L.tm_mesh_tfi_block(self._h, block, *[_dp(a) for a in arrs]))

    def tfi_block_resident(self, block: int):
        check(self._L.tm_mesh_tfi_block_resident(self._h, block))

    def upload_block(self, block: int, xy: np.ndarray):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        check(self._L.tm_mesh_upload_block(self._h, block, _dp(xy)))

    def download_block(self, block: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        ni, nj = C.c_uint64(), C.c_uint64()
        check(self._L.tm_mesh_block_size(self._h, block, C.byref(ni), C.byref(nj)))
        if out is None:
            out = np.empty((ni.value, nj.value, 2), dtype=np.float64)
        check(self._L.tm_mesh_download_block(self._h, block, _dp(out)))
        return out

    def download_block_async(self, block: int, out: np.ndarray) -> None:
        """Starts the copy-back of a block into ``out`` (pinned host memory) and returns; the mesh may be reused at once
        (``tm_mesh_download_block_async``).  ``out`` is valid after :meth:`download_wait`."""
        check(self._L.tm_mesh_download_block_async(self._h, block, _dp(out)))

    def download_wait(self) -> None:
        check(self._L.tm_mesh_download_wait(self._h))

    def download(self):
        """Copies all blocks held by this process back into ``self.mesh`` (in place, like smooth.zig:139-153)."""
        for k, b in enumerate(self.mesh.blocks):
            if self.owns(k):
                self.download_block(k, b.points)
        return self.mesh

    def set_white_groups(self, block_pairs):
        """One (A, B) pair of O-grid half blocks per cut of a batch (``tm_mesh_set_white_groups``)."""
        flat = [int(v) for pair in block_pairs for v in pair]
        arr = (C.c_uint64 * max(len(flat), 1))(*flat)
        check(self._L.tm_mesh_set_white_groups(self._h, arr, len(flat) // 2))

    def begin_smoothing(self, solver: Optional[CudaSolver] = None, control_function=None):
        self._opts = make_options(0, solver, control_function)
        check(self._L.tm_mesh_begin_smoothing(self._h, C.byref(self._opts)))

    def smooth(self, iterations: int, solver: Optional[CudaSolver] = None, control_function=None) -> dict:
        opts = make_options(iterations, solver, control_function)
        st = TmSmoothStats()
        check(self._L.tm_mesh_smooth(self._h, C.byref(opts), C.byref(st)))
        return st.as_dict()

    def synchronize(self):
        check(self._L.tm_mesh_synchronize(self._h))

    def control_function(self, block: int) -> np.ndarray:
        ni, nj = C.c_uint64(), C.c_uint64()
        check(self._L.tm_mesh_block_size(self._h, block, C.byref(ni), C.byref(nj)))
        out = np.empty((ni.value, nj.value, 2), dtype=np.float64)
        check(self._L.tm_mesh_download_control_function(self._h, block, _dp(out)))
        return out

    def block_soa(self, block: int, field: str = "coordinates"):
        """Structured output of one block (``tm_mesh_download_block_soa``): two flat arrays with i fastest, the buffers
        ``cgns.write`` hands to ``cg_coord_write`` / ``cg_field_write`` (cgns.zig:69-101, 110-161)."""
        ni, nj = C.c_uint64(), C.c_uint64()
        check(self._L.tm_mesh_block_size(self._h, block, C.byref(ni), C.byref(nj)))
        x, y = np.empty(ni.value * nj.value), np.empty(ni.value * nj.value)
        check(self._L.tm_mesh_download_block_soa(self._h, block, {"coordinates": 0, "control_function": 1}[field], _dp(x), _dp(y)))
        return x, y

    def write_plot3d(self, grid_path: str, function_path: Optional[str] = None):
        """Multi-block PLOT3D grid file (and, optionally, the control function as a function file) of the blocks this
        process holds -- the structured output step right after the path (``cgns.write``, cgns.zig:26-168)."""
        check(self._L.tm_mesh_write_plot3d(self._h, os.fsencode(grid_path), os.fsencode(function_path) if function_path else None))

    def viewer_buffers(self):
        """f32 point buffer, (x_min, x_max, y_min, y_max) and wireframe line indices built on the device
        (``tm_mesh_viewer_buffers``; gui/lib.zig:227-318)."""
        n_pts, n_idx = C.c_uint64(), C.c_uint64()
        check(self._L.tm_mesh_viewer_sizes(self._h, C.byref(n_pts), C.byref(n_idx)))
        pts = np.empty(2 * n_pts.value, dtype=np.float32)
        rng = np.empty(4, dtype=np.float32)
        idx = np.empty(n_idx.value, dtype=np.uint32)
        fp = C.POINTER(C.c_float)
        check(self._L.tm_mesh_viewer_buffers(self._h, pts.ctypes.data_as(fp), rng.ctypes.data_as(fp), idx.ctypes.data_as(C.POINTER(C.c_uint32))))
        return pts, rng, idx

    def boundary_kinds(self, block: int) -> np.ndarray:
        ni, nj = C.c_uint64(), C.c_uint64()
        check(self._L.tm_mesh_block_size(self._h, block, C.byref(ni), C.byref(nj)))
        out = np.empty(2 * (ni.value + nj.value - 2), dtype=np.uint8)
        check(self._L.tm_mesh_download_boundary_kinds(self._h, block, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    @property
    def component_count(self) -> int:
        """Independent systems of the mesh: connected components of the block graph (the cuts of a batch)."""
        return int(self._L.tm_mesh_component_count(self._h))

    def component_of_block(self, block: int) -> int:
        c = C.c_uint64()
        check(self._L.tm_mesh_component_of_block(self._h, block, C.byref(c)))
        return int(c.value)

    def component_stats(self, component: int) -> dict:
        """Record of the inner solves of ``component`` in the last outer iteration (``tm_mesh_component_stats``)."""
        st = _lib.TmComponentStats()
        check(self._L.tm_mesh_component_stats(self._h, component, C.byref(st)))
        return st.as_dict()

    def block_device_ptr(self, block: int) -> int:
        return int(self._L.tm_mesh_block_device_ptr(self._h, block) or 0)


def dist_unique_id() -> bytes:
    """NCCL unique id (rank 0 creates it and broadcasts the bytes to the other ranks)."""
    buf = (C.c_uint8 * _lib.TM_UNIQUE_ID_BYTES)()
    check(_lib.load().tm_dist_get_unique_id(buf))
    return bytes(buf)


def dist_plan(mesh, owner, rank: int, n_ranks: int) -> dict:
    """Host-only partition plan of one rank (``tm_dist_plan``): sizes plus the ghost / send id lists per peer."""
    L = _lib.load()
    cm = _CMesh(mesh, with_coords=False)
    own = (C.c_int32 * cm.nb)(*[int(o) for o in owner])
    info = _lib.TmDistPlanInfo()
    counts = (C.c_int64 * (2 * n_ranks))()
    check(L.tm_dist_plan(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, own, rank, n_ranks, C.byref(info), None, None, counts))
    ghost = (C.c_int64 * max(int(info.n_ghost), 1))()
    send = (C.c_int64 * max(int(info.n_send), 1))()
    check(L.tm_dist_plan(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, own, rank, n_ranks, C.byref(info), ghost, send, counts))
    out = info.as_dict()
    g, s_, ghost_ids, send_ids = 0, 0, [], []
    for p in range(n_ranks):
        ng, ns = int(counts[2 * p]), int(counts[2 * p + 1])
        ghost_ids.append(np.array(ghost[g:g + ng], dtype=np.int64))
        send_ids.append(np.array(send[s_:s_ + ns], dtype=np.int64))
        g, s_ = g + ng, s_ + ns
    out["ghost_ids"], out["send_ids"] = ghost_ids, send_ids
    return out


def mg_plan(mesh, cell_size=None, max_levels: int = 32):
    """Host-only plan of the multigrid hierarchy (``tm_mg_plan``): list over levels of the (ni, nj) of every block."""
    L = _lib.load()
    cm = _CMesh(mesh, with_coords=False)
    n = C.c_uint64()
    sizes = (C.c_uint64 * (max_levels * max(cm.nb, 1) * 2))()
    h = None
    if cell_size is not None:
        hs = np.ascontiguousarray(cell_size, dtype=np.float64).ravel()
        assert hs.size == 2 * cm.nb
        h = _dp(hs)
    check(L.tm_mg_plan(cm.blocks, cm.nb, cm.conns, cm.nc, cm.bcs, cm.nbc, h, max_levels, C.byref(n), sizes))
    return [[(int(sizes[(l * cm.nb + b) * 2]), int(sizes[(l * cm.nb + b) * 2 + 1])) for b in range(cm.nb)] for l in range(min(int(n.value), max_levels))]


def stream_plan(ni: int, nj: int, sweeps: int):
    """Host-only plan of the streamed ``tm_smooth_mesh`` (``tm_smooth_stream_plan``): ``None`` when the block is smoothed
    resident, else ``(window_rows, window_first[k], owned_first[k] ... owned_first[K])``."""
    L = _lib.load()
    n, w = C.c_uint64(), C.c_uint64()
    first, owned = (C.c_uint64 * 8)(), (C.c_uint64 * 9)()
    check(L.tm_smooth_stream_plan(ni, nj, sweeps, C.byref(n), C.byref(w), first, owned))
    if n.value == 0:
        return None
    return int(w.value), [int(v) for v in first[:n.value]], [int(v) for v in owned[:n.value + 1]]


def kernel_launch_count() -> int:
    return int(_lib.load().tm_kernel_launch_count())


def device_info(device: int = -1) -> dict:
    name = C.create_string_buffer(256)
    sms, mem = C.c_int(), C.c_uint64()
    check(_lib.load().tm_device_info(device, name, 256, C.byref(sms), C.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sms.value, "global_mem_bytes": mem.value}


def read_plot3d(path: str, n_vars: Optional[int] = None):
    """Reads a file written by :meth:`DeviceMesh.write_plot3d`: a list of (ni, nj, 2) arrays (x, y of a grid file, or P, Q
    of a function file when ``n_vars`` is given)."""
    with open(path, "rb") as f:
        nb = int(np.fromfile(f, dtype=np.int32, count=1)[0])
        width = 2 if n_vars is None else 3
        dims = np.fromfile(f, dtype=np.int32, count=width * nb).reshape(nb, width)
        out = []
        for ni, nj in dims[:, :2]:
            a = np.fromfile(f, dtype=np.float64, count=2 * int(ni) * int(nj)).reshape(2, int(nj), int(ni))
            out.append(np.ascontiguousarray(a.transpose(2, 1, 0)))
        assert f.read(1) == b""
    return out
