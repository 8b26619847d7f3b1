import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
n_bi, n_bj, ni, nj = 8, 8, 129, 65
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
    return out, st
ref, st = solve(120)
print("ref last", st["last_max_update"])
for c in ():
    cur, st = solve(c)
    print(f"cycles {c}: last_max_update {st['last_max_update']:.2e}")
    for bj in reversed(range(n_bj)):
        row = []
        for bi in range(n_bi):
            e = np.abs(cur[bi*n_bj+bj] - ref[bi*n_bj+bj]).max(axis=2)
            idx = np.unravel_index(e.argmax(), e.shape)
            row.append(f"{e.max():.1e}@({idx[0]:3d},{idx[1]:2d})")
        print("  bj", bj, " ".join(row))
np.set_printoptions(linewidth=250, precision=2)
prev = None
for c in (12, 13, 14):
    cur, st = solve(c)
    line = []
    for bi in range(n_bi):
        e = cur[bi*n_bj+6] - ref[bi*n_bj+6]
        line += [e[0, 32, 1], e[64, 32, 1]]
    print(f"cycles {c} ey along i (bj=6, j=32):", np.array(line))
    line = []
    for bi in range(n_bi):
        e = cur[bi*n_bj+6] - ref[bi*n_bj+6]
        line += [e[0, 32, 0], e[64, 32, 0]]
    print(f"cycles {c} ex along i (bj=6, j=32):", np.array(line))
