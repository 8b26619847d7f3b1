"""Structured (SoA) output of an 8192^2 block: transpose on the device + D2H into pinned memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from turbomesh_b200 import smoothing, synthetic
spec = synthetic.single_block(8192, 8192)
dm = smoothing.DeviceMesh(spec, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
import ctypes as C
n = 8192 * 8192
x = torch.empty(n, dtype=torch.float64, pin_memory=True); y = torch.empty(n, dtype=torch.float64, pin_memory=True)
dp = C.POINTER(C.c_double)
for _ in range(3):
    t0 = time.perf_counter()
    smoothing.check(dm._L.tm_mesh_download_block_soa(dm._h, 0, 0, C.cast(x.data_ptr(), dp), C.cast(y.data_ptr(), dp)))
    t1 = time.perf_counter()
    print(f"SoA download 8192^2 (transpose on device + 1.07 GB D2H, pinned): {(t1-t0)*1e3:.1f} ms")
dm.close()
