"""Step time of the one-shot host-buffer entry points (tm_tfi_block + tm_smooth_mesh) after different amounts of prior GPU work."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
bench.pin_to_gpu_numa_node(0)
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Block2d, Mesh
spec = synthetic.single_block(8192, 8192)
stream = torch.cuda.Stream()
dm = smoothing.DeviceMesh(spec, device=0, stream=stream.cuda_stream, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=100, omega=0.9, device=0)
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
if mode in ("full", "relax"):
    for _ in range(6):
        dm.tfi_block_resident(0); dm.begin_smoothing(solver); dm.smooth(1, solver)
if mode == "full":
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-10, device=0)
    for _ in range(2):
        dm.tfi_block_resident(0); dm.begin_smoothing(mg); dm.smooth(100, mg)
dm.synchronize()
b = spec.blocks[0]
pinned = torch.empty((8192, 8192, 2), dtype=torch.float64, pin_memory=True)
host = pinned.numpy()
edges = b.edge_args()
mesh = Mesh([Block2d.__new__(Block2d)], ["block"], [], [])
mesh.blocks[0].points = host
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
for it in range(int(os.environ.get('PROBE_ITERS', '4'))):
    t1, _ = T(lambda: smoothing.tfi_block(*edges, out=host))
    t2, _ = T(lambda: smoothing.smooth_mesh(mesh, 1, solver))
    print(f"[{mode}] tfi_block {t1:.1f} ms, smooth_mesh {t2:.1f} ms")
