"""turbomesh_b200 -- B200 (sm_100a) back-end for turbomesh's hot path: 2D TFI + multi-block elliptic smoothing.

Host-side mirror of the part of the reference's ``src/core`` block/mesh API the path sits behind (``discrete``, ``boundary``,
``clustering``, ``smoothing``) whose two hot calls -- ``Block2d.init`` (TFI) and ``smoothing.mesh`` -- go through the C ABI of
``include/turbomesh_gpu.h`` to hand-written CUDA kernels; ``blocking`` is the O4H template batched over cuts on the device (the
caller right in front of the path); ``synthetic`` holds the workloads of the named shapes.
"""
from . import boundary, clustering, discrete  # noqa: F401

__all__ = ["blocking", "boundary", "clustering", "discrete", "smoothing", "synthetic"]
