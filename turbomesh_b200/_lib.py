"""ctypes binding of the C ABI (``include/turbomesh_gpu.h``) -- the only way Python reaches the GPU code.

The shared library is built in-tree by ``turbomesh_b200/build.py`` (nvcc, sm_100a).  Loading fails loudly
when it is missing; there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libturbomesh_gpu.so")

TM_OK = 0
TM_ERR_INVALID_ARGUMENT, TM_ERR_NO_DEVICE, TM_ERR_CUDA, TM_ERR_TOPOLOGY = -1, -2, -3, -4
TM_ERR_UNSUPPORTED, TM_ERR_NOT_CONVERGED, TM_ERR_OUT_OF_MEMORY = -5, -6, -7
TM_SOLVER_PICARD_BICGSTAB, TM_SOLVER_RELAX, TM_SOLVER_FAS_MULTIGRID = 0, 1, 2
TM_CF_LAPLACE, TM_CF_WHITE = 0, 1


class TmBlock(C.Structure):
    _fields_ = [("ni", C.c_uint64), ("nj", C.c_uint64), ("xy", C.POINTER(C.c_double))]


class TmRange(C.Structure):
    _fields_ = [("block", C.c_uint64), ("side", C.c_uint32), ("_pad", C.c_uint32), ("start", C.c_uint64), ("end", C.c_uint64)]


class TmConnection(C.Structure):
    _fields_ = [("ranges", TmRange * 2), ("has_periodicity", C.c_int32), ("_pad", C.c_int32), ("periodicity", C.c_double * 2)]


class TmCondition(C.Structure):
    _fields_ = [("range", TmRange), ("kind", C.c_uint32), ("_pad", C.c_uint32)]


class TmSmoothOptions(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("solver", C.c_uint32), ("iterations", C.c_uint64), ("control_function", C.c_uint32),
                ("fail_on_no_convergence", C.c_uint32), ("white_ds_target", C.c_double), ("white_theta_target", C.c_double),
                ("rtol", C.c_double), ("atol", C.c_double), ("max_inner_iterations", C.c_uint64), ("omega", C.c_double),
                ("sweeps_per_iteration", C.c_uint64), ("stop_max_update", C.c_double), ("device", C.c_int32), ("inner_refinement_cycles", C.c_int32)]


class TmSmoothStats(C.Structure):
    _fields_ = [("outer_iterations", C.c_uint64), ("inner_iterations", C.c_uint64), ("operator_applications", C.c_uint64), ("nodes", C.c_uint64),
                ("last_sumsq_x", C.c_double), ("last_sumsq_y", C.c_double), ("last_residual", C.c_double), ("last_max_update", C.c_double),
                ("last_inner_residual", C.c_double), ("gpu_seconds", C.c_double), ("converged", C.c_int32), ("streamed_chunks", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("_")}


class TmDistPlanInfo(C.Structure):
    _fields_ = [("n_own", C.c_uint64), ("n_ghost", C.c_uint64), ("n_synth", C.c_uint64), ("n_send", C.c_uint64),
                ("n_smoothed", C.c_uint64), ("n_junction", C.c_uint64), ("n_sliding", C.c_uint64), ("n_slaves", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


TM_UNIQUE_ID_BYTES = 128


class TmSpline(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("params", C.POINTER(C.c_double)), ("points", C.POINTER(C.c_double)),
                ("second_derivs_x", C.POINTER(C.c_double)), ("second_derivs_y", C.POINTER(C.c_double)),
                ("n_samples", C.c_uint64), ("sample_arc", C.POINTER(C.c_double)), ("total_length", C.c_double)]


class TmEdgeJob(C.Structure):
    _fields_ = [("n", C.c_uint64), ("curve_kind", C.c_uint32), ("clustering_kind", C.c_uint32),
                ("line_start", C.c_double * 2), ("line_end", C.c_double * 2), ("spline", C.POINTER(TmSpline)),
                ("alpha", C.c_double), ("beta", C.c_double), ("delta_s", C.c_double),
                ("points", C.c_void_p), ("clustering", C.c_void_p)]   # double*: plain addresses, set from numpy without a cast per job


class TmComponentStats(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("iterations", C.c_uint64 * 2), ("tolerance", C.c_double * 2), ("norm_b", C.c_double * 2),
                ("norm_r", C.c_double * 2), ("status", C.c_int32 * 2), ("operator_applications", C.c_uint64), ("restarts", C.c_uint64)]

    def as_dict(self):
        return {"nodes": int(self.nodes), "iterations": tuple(self.iterations), "tolerance": tuple(self.tolerance), "norm_b": tuple(self.norm_b),
                "norm_r": tuple(self.norm_r), "status": tuple(self.status), "operator_applications": int(self.operator_applications),
                "restarts": int(self.restarts)}


class TmSplineFitJob(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("points", C.POINTER(C.c_double)), ("n_samples", C.c_uint64), ("params", C.POINTER(C.c_double)),
                ("second_derivs_x", C.POINTER(C.c_double)), ("second_derivs_y", C.POINTER(C.c_double)), ("sample_arc", C.POINTER(C.c_double)),
                ("total_length", C.POINTER(C.c_double))]


class TmEdgeView(C.Structure):
    _fields_ = [("points", C.c_void_p), ("clustering", C.c_void_p), ("n", C.c_uint64), ("start", C.c_uint64), ("end", C.c_uint64)]


class TmCombineJob(C.Structure):
    _fields_ = [("views", C.c_void_p), ("n_views", C.c_uint64), ("points", C.c_void_p), ("clustering", C.c_void_p)]


class TmProjectJob(C.Structure):
    _fields_ = [("points", C.c_void_p), ("n", C.c_uint64), ("distance", C.c_double), ("out", C.c_void_p)]


class TurbomeshGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"turbomesh_gpu error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load():
    """Loads ``libturbomesh_gpu.so`` and declares every entry point of ``include/turbomesh_gpu.h``."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m turbomesh_b200.build` (nvcc, sm_100a). "
                          "turbomesh_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.tm_last_error.restype = C.c_char_p
    L.tm_abi_version.restype = C.c_int
    L.tm_kernel_launch_count.restype = C.c_uint64
    L.tm_release_cached_memory.restype = None
    L.tm_device_info.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    L.tm_smooth_options_default.argtypes = [C.POINTER(TmSmoothOptions)]
    L.tm_smooth_options_default.restype = None
    L.tm_tfi_block.argtypes = [C.c_uint64, C.c_uint64] + [dp] * 9
    L.tm_smooth_mesh.argtypes = [C.POINTER(TmBlock), C.c_size_t, C.POINTER(TmConnection), C.c_size_t, C.POINTER(TmCondition), C.c_size_t,
                                 C.POINTER(TmSmoothOptions), C.POINTER(TmSmoothStats)]
    L.tm_mesh_create.argtypes = [C.POINTER(TmBlock), C.c_size_t, C.POINTER(TmConnection), C.c_size_t, C.POINTER(TmCondition), C.c_size_t,
                                 C.c_int, vp, C.POINTER(vp)]
    L.tm_dist_get_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.tm_mesh_create_distributed.argtypes = [C.POINTER(TmBlock), C.c_size_t, C.POINTER(TmConnection), C.c_size_t, C.POINTER(TmCondition), C.c_size_t,
                                             C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int, vp, C.POINTER(vp)]
    L.tm_mesh_local_node_count.argtypes = [vp]
    L.tm_mesh_local_node_count.restype = C.c_uint64
    L.tm_mesh_halo_path.argtypes = [vp]
    L.tm_mesh_halo_path.restype = C.c_int
    L.tm_dist_plan.argtypes = [C.POINTER(TmBlock), C.c_size_t, C.POINTER(TmConnection), C.c_size_t, C.POINTER(TmCondition), C.c_size_t,
                               C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(TmDistPlanInfo), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                               C.POINTER(C.c_int64)]
    L.tm_mg_plan.argtypes = [C.POINTER(TmBlock), C.c_size_t, C.POINTER(TmConnection), C.c_size_t, C.POINTER(TmCondition), C.c_size_t, dp, C.c_size_t,
                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tm_edges_discretize.argtypes = [C.POINTER(TmEdgeJob), C.c_size_t, C.c_int]
    L.tm_splines_fit.argtypes = [C.POINTER(TmSplineFitJob), C.c_size_t, C.c_int]
    L.tm_edges_combine.argtypes = [C.POINTER(TmCombineJob), C.c_size_t, C.c_int]
    L.tm_edges_project_normal.argtypes = [C.POINTER(TmProjectJob), C.c_size_t, C.c_int]
    L.tm_smooth_stream_plan.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64] + [C.POINTER(C.c_uint64)] * 4
    L.tm_mesh_destroy.argtypes = [vp]
    L.tm_mesh_destroy.restype = None
    L.tm_mesh_upload_block.argtypes = [vp, C.c_size_t, dp]
    L.tm_mesh_download_block.argtypes = [vp, C.c_size_t, dp]
    L.tm_mesh_download_block_async.argtypes = [vp, C.c_size_t, dp]
    L.tm_mesh_download_wait.argtypes = [vp]
    L.tm_mesh_tfi_block.argtypes = [vp, C.c_size_t] + [dp] * 8
    L.tm_mesh_tfi_block_resident.argtypes = [vp, C.c_size_t]
    L.tm_mesh_set_white_groups.argtypes = [vp, C.POINTER(C.c_uint64), C.c_size_t]
    L.tm_mesh_begin_smoothing.argtypes = [vp, C.POINTER(TmSmoothOptions)]
    L.tm_mesh_smooth.argtypes = [vp, C.POINTER(TmSmoothOptions), C.POINTER(TmSmoothStats)]
    L.tm_mesh_synchronize.argtypes = [vp]
    L.tm_mesh_component_count.argtypes = [vp]
    L.tm_mesh_component_count.restype = C.c_uint64
    L.tm_mesh_component_of_block.argtypes = [vp, C.c_size_t, C.POINTER(C.c_uint64)]
    L.tm_mesh_component_stats.argtypes = [vp, C.c_size_t, C.POINTER(TmComponentStats)]
    L.tm_mesh_block_count.argtypes = [vp]
    L.tm_mesh_block_count.restype = C.c_uint64
    L.tm_mesh_node_count.argtypes = [vp]
    L.tm_mesh_node_count.restype = C.c_uint64
    L.tm_mesh_block_size.argtypes = [vp, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tm_mesh_block_device_ptr.argtypes = [vp, C.c_size_t]
    L.tm_mesh_block_device_ptr.restype = vp
    L.tm_mesh_download_control_function.argtypes = [vp, C.c_size_t, dp]
    L.tm_mesh_download_boundary_kinds.argtypes = [vp, C.c_size_t, C.POINTER(C.c_uint8)]
    L.tm_mesh_download_block_soa.argtypes = [vp, C.c_size_t, C.c_int, dp, dp]
    L.tm_mesh_write_plot3d.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.tm_mesh_viewer_sizes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tm_mesh_viewer_buffers.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    _lib = L
    return L


def check(rc: int):
    if rc != TM_OK:
        raise TurbomeshGpuError(rc, load().tm_last_error().decode(errors="replace"))
