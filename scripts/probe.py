"""Quick device-time probe of the TFI and sweep kernels (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
print(smoothing.device_info())
spec = synthetic.single_block(n, n)
t0 = time.time()
dm = smoothing.DeviceMesh(spec, upload=False)
print("create", time.time() - t0)
b = spec.blocks[0]
t0 = time.time(); dm.tfi_block(0, *b.edge_args()); dm.synchronize(); print("tfi first (with edge upload)", time.time() - t0)
for _ in range(3):
    t0 = time.time(); dm.tfi_block_resident(0); dm.synchronize(); dt = time.time() - t0
    print("tfi resident %.3f ms  -> %.1f GB/s (16 B/node)" % (dt * 1e3, n * n * 16 / dt / 1e9))
sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=1.0)
dm.begin_smoothing(sol)
for _ in range(3):
    st = dm.smooth(1, sol)
    per = st["gpu_seconds"] / sweeps
    print("relax sweep %.3f ms -> %.1f G node-updates/s, %.1f GB/s (32 B/node-update); max_update %.3e" % (per * 1e3, n * n / per / 1e9, n * n * 32 / per / 1e9, st["last_max_update"]))
if len(sys.argv) > 3:
    sol2 = smoothing.CudaSolver(method="picard_bicgstab", rtol=1e-6, atol=1e-8, max_inner_iterations=int(sys.argv[3]))
    st = dm.smooth(1, sol2)
    print("picard step", st)
