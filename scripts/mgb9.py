import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
n = 129
n_bi, n_bj = 1, 8
spec = synthetic.cascade(n_bi, n_bj, 2 * n - 1, n, length=0.125)
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
ref, st = solve(150)
print("ref last update", st["last_max_update"])
for c in (10, 11):
    cur, _ = solve(c)
    print(f"cycles {c}")
    for b in range(n_bj):
        e = cur[b] - ref[b]
        ex, ey = e[..., 0], e[..., 1]
        ni, nj = ex.shape
        print(f"  block {b}: max|ex| {np.abs(ex).max():.2e} max|ey| {np.abs(ey).max():.2e};  ey at i=0: j=0 {ey[0,0]:+.2e} j=mid {ey[0,nj//2]:+.2e} j=end {ey[0,-1]:+.2e} | i=mid: {ey[ni//2,0]:+.2e} {ey[ni//2,nj//2]:+.2e} {ey[ni//2,-1]:+.2e} | ex i=mid {ex[ni//2,0]:+.2e} {ex[ni//2,nj//2]:+.2e} {ex[ni//2,-1]:+.2e}")
cur, _ = solve(11)
e = cur[3] - ref[3]
ey = e[..., 1]
ni, nj = ey.shape
idx = np.unravel_index(np.abs(ey).argmax(), ey.shape)
print("block 3 argmax", idx, "of", (ni, nj))
np.set_printoptions(linewidth=250, precision=2)
for (i0, j0) in ((0, nj - 10), (ni - 10, nj - 10), (0, 0)):
    print(f"ey[{i0}:{i0+10}, {j0}:{j0+10}] =")
    print(ey[i0:i0 + 10, j0:j0 + 10])
print("ey along i at j=nj-2, every 16:", ey[::16, nj - 2])
print("ey along i at j=nj-9, every 16:", ey[::16, nj - 9])
