"""turbomesh_b200 -- B200 (sm_100a) back-end for turbomesh's hot path: 2D TFI + multi-block elliptic smoothing.

Host-side mirror of the reference's ``src/core`` block/mesh API (``discrete``, ``boundary``, ``clustering``,
``spline``, ``geometry``, ``templates``, ``input``, ``smoothing``) whose two hot calls -- ``Block2d.init`` (TFI) and
``smoothing.mesh`` -- go through the C ABI of ``include/turbomesh_gpu.h`` to hand-written CUDA kernels.
"""
from . import boundary, clustering, discrete, geometry, spline  # noqa: F401

__all__ = ["boundary", "clustering", "discrete", "geometry", "spline", "templates", "input", "smoothing", "synthetic"]
