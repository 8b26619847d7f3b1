"""FAS multigrid vs relaxation / Picard on a single block: convergence history and time to converged."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 2
omega = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-10
spec = synthetic.single_block(n, n)
dm = smoothing.DeviceMesh(spec, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
sol = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=omega)
dm.begin_smoothing(sol)
tot = 0.0
for c in range(60):
    st = dm.smooth(1, sol)
    tot += st["gpu_seconds"]
    print(f"cycle {c:2d}: max_update {st['last_max_update']:.3e}  cycle time {st['gpu_seconds']*1e3:.3f} ms  fine-equivalent ops {st['operator_applications']}")
    if st["last_max_update"] <= tol:
        break
print(f"n={n}: converged to {tol:g} in {c+1} cycles, {tot*1e3:.2f} ms of GPU time")
if n <= 257:
    from oracle import oracle as orc
    mg = dm.download_block(0)
    ref = synthetic.materialize(spec, orc.tfi)
    orc.smooth_mesh(ref, 40, orc.tight_options())
    print("max |mg - oracle(40 tight Picard its)| = %.3e" % np.abs(mg - ref.blocks[0].points).max())
