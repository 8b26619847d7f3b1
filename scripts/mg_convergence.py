"""Convergence histories of the (Anderson-accelerated) multigrid on 8x8 / 2x8 / 4x4 block cascades; TM_MG_AA=0 for the plain cycle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
def hist(spec, nu=3, omega=0.8, cycles=40, tol=1e-11):
    dm = smoothing.DeviceMesh(spec, upload=False)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=omega)
    dm.begin_smoothing(mg)
    h = []
    for c in range(cycles):
        st = dm.smooth(1, mg)
        h.append(st["last_max_update"])
        if h[-1] < tol:
            break
    dm.close()
    return h
n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
for name, mk in [("8x8", lambda: synthetic.cascade(8, 8, (n + 1) // 2, (n + 3) // 4)), ("2x8", lambda: synthetic.cascade(2, 8, 2 * n - 1, n, length=0.25, ay=0.015 / 4)),
                 ("4x4", lambda: synthetic.cascade(4, 4, n, (n + 1) // 2))]:
    h = hist(mk())
    print(f"AA={os.environ.get('TM_MG_AA')} n={n} {name}: {len(h)} cycles to {h[-1]:.1e}:", " ".join(f"{v:.1e}" for v in h), flush=True)
