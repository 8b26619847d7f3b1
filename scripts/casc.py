"""Where does the cascade sweep lose time vs the single block?  (a) cascade, (b) same blocks without any connection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Mesh
sweeps = 200
for name in ("cascade", "unconnected", "single"):
    if name == "single":
        spec = synthetic.single_block(8192, 8192)
    else:
        spec = synthetic.cascade(1, 8, 4096, 2048)
        if name == "unconnected":
            spec = Mesh(blocks=spec.blocks, names=spec.names, connections=[], boundary_conditions=[])
    dm = smoothing.DeviceMesh(spec, upload=False)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=0.9)
    dm.begin_smoothing(sol)
    best = 1e9
    for _ in range(3):
        st = dm.smooth(1, sol); best = min(best, st["gpu_seconds"] / sweeps)
    n = dm.node_count
    print(f"{name:12s} {best*1e3:.4f} ms/sweep  {n*32/best/1e9:.1f} GB/s")
    dm.close()
