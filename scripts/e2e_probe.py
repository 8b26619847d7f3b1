import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
print("affinity:", bench.pin_to_gpu_numa_node(0))
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Block2d, Mesh
spec = synthetic.single_block(8192, 8192)
b = spec.blocks[0]
pinned = torch.empty((8192, 8192, 2), dtype=torch.float64, pin_memory=True)
host = pinned.numpy()
edges = b.edge_args()
mesh = Mesh([Block2d.__new__(Block2d)], ["block"], [], [])
mesh.blocks[0].points = host
solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=100, omega=0.9)
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
for it in range(3):
    t1, _ = T(lambda: smoothing.tfi_block(*edges, out=host))
    t2, _ = T(lambda: smoothing.smooth_mesh(mesh, 1, solver))
    print(f"tfi_block {t1:.1f} ms, smooth_mesh {t2:.1f} ms")
# pieces
t, dm = T(lambda: smoothing.DeviceMesh(mesh, upload=False)); print(f"create (no upload) {t:.1f} ms")
t, _ = T(lambda: dm.upload_block(0, host)); print(f"upload 1.07 GB {t:.1f} ms")
t, _ = T(lambda: dm.begin_smoothing(solver)); print(f"begin {t:.1f} ms")
t, _ = T(lambda: dm.smooth(1, solver)); print(f"100 sweeps {t:.1f} ms")
t, _ = T(lambda: dm.download_block(0, host)); print(f"download 1.07 GB {t:.1f} ms")
t, _ = T(lambda: dm.close()); print(f"destroy {t:.1f} ms")
dev = torch.empty_like(pinned, device="cuda")
for _ in range(2):
    t, _ = T(lambda: dev.copy_(pinned, non_blocking=True)); print(f"torch h2d {t:.1f} ms")
    t, _ = T(lambda: pinned.copy_(dev, non_blocking=True)); print(f"torch d2h {t:.1f} ms")
