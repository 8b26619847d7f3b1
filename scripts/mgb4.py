import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
os.environ["TM_MG_NESTED"] = "1"
spec = synthetic.cascade(1, 1, 129, 65)
def solve(cycles):
    dm = smoothing.DeviceMesh(spec, upload=False)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
    dm.begin_smoothing(mg)
    st = dm.smooth(cycles, mg)
    out = dm.download_block(0)
    dm.close()
    return out
ref = solve(80)
for c in (6, 7, 8):
    e = solve(c) - ref
    ex, ey = np.abs(e[..., 0]), np.abs(e[..., 1])
    print(f"cycles {c}: max|ex| {ex.max():.2e} max|ey| {ey.max():.2e}")
    print("  |ey| max per i (0,1,2,4,8,16,32,64,96,120,126,127,128):", " ".join(f"{ey[i].max():.1e}" for i in (0,1,2,4,8,16,32,64,96,120,126,127,128)))
    print("  |ex| max per i:", " ".join(f"{ex[i].max():.1e}" for i in (0,1,2,4,8,16,32,64,96,120,126,127,128)))
    print("  |ey| row 0 at j (0,1,2,4,8,16,32,48,60,63,64):", " ".join(f"{e[0,j,1]:+.1e}" for j in (0,1,2,4,8,16,32,48,60,63,64)))
    print("  |ey| row 4 at j:", " ".join(f"{e[4,j,1]:+.1e}" for j in (0,1,2,4,8,16,32,48,60,63,64)))
