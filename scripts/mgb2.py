import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
n_bi, n_bj, ni, nj = 4, 4, 129, 65
spec = synthetic.cascade(n_bi, n_bj, ni, nj)
dm = smoothing.DeviceMesh(spec, upload=False)
for k, b in enumerate(spec.blocks):
    dm.tfi_block(k, *b.edge_args())
mode = sys.argv[1] if len(sys.argv) > 1 else "mg"
omega = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=omega)
rl = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=omega)
dm.begin_smoothing(mg)
if mode == "mg":
    for c in range(30):
        st = dm.smooth(1, mg)
    print("mg last", st["last_max_update"])
else:
    rl100 = smoothing.CudaSolver(method="relax", sweeps_per_iteration=500, omega=omega)
    for c in range(20):
        st = dm.smooth(1, rl100)
        print("relax", (c+1)*500, st["last_max_update"])
prev = [dm.download_block(k) for k in range(len(spec.blocks))]
for sweep in range(6):
    st = dm.smooth(1, rl)
    cur = [dm.download_block(k) for k in range(len(spec.blocks))]
    worst = max(range(len(cur)), key=lambda k: np.abs(cur[k]-prev[k]).max())
    d = cur[worst]-prev[worst]
    idx = np.unravel_index(np.abs(d).argmax(), d.shape)
    print(f"sweep {sweep}: max update {st['last_max_update']:.3e} block {worst} (bi={worst//n_bj}, bj={worst%n_bj}) node {idx} d={d[idx[0], idx[1]]}")
    prev = cur
