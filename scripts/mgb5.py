import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
n_bi, n_bj, ni, nj = 4, 4, 65, 33
spec = synthetic.cascade(n_bi, n_bj, ni, nj)
def solve(cycles):
    dm = smoothing.DeviceMesh(spec, upload=False)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
    dm.begin_smoothing(mg)
    st = dm.smooth(cycles, mg)
    out = [dm.download_block(k) for k in range(len(spec.blocks))]
    dm.close()
    return out
ref = solve(80)
for c in (8, 9):
    cur = solve(c)
    print(f"cycles {c}")
    for bi in range(n_bi):
        row = []
        for bj in range(n_bj):
            e = np.abs(cur[bi*n_bj+bj] - ref[bi*n_bj+bj]).max(axis=2)
            idx = np.unravel_index(e.argmax(), e.shape)
            row.append(f"{e.max():.1e}@({idx[0]:2d},{idx[1]:2d})")
        print("  bi", bi, " ".join(row))
