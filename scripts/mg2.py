import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
ni, nj, nu, omega = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
spec = synthetic.single_block(ni, nj)
dm = smoothing.DeviceMesh(spec, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
sol = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=omega)
dm.begin_smoothing(sol)
prev = None
for c in range(int(sys.argv[5]) if len(sys.argv) > 5 else 30):
    st = dm.smooth(1, sol)
    print(c, "%.3e" % st["last_max_update"], ("ratio %.3f" % (st["last_max_update"] / prev)) if prev else "")
    prev = st["last_max_update"]
