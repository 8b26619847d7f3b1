"""Sweep time of the cascade workload for two block sizes (alignment check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
sweeps = 200
for (ni, nj) in ((4096, 2048), (4097, 2049), (4097, 2049)):
    spec = synthetic.cascade(1, 8, ni, nj, length=0.125, ay=0.015 / 8)
    dm = smoothing.DeviceMesh(spec, upload=False)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=0.9)
    dm.begin_smoothing(sol)
    best = 1e9
    for _ in range(3):
        st = dm.smooth(1, sol); best = min(best, st["gpu_seconds"] / sweeps)
    n = dm.node_count
    print(f"{ni}x{nj}: {best*1e3:.4f} ms/sweep  {n*32/best/1e9:.1f} GB/s")
    dm.close()
