import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
def rate(spec, nu, omega, cycles=14):
    dm = smoothing.DeviceMesh(spec, upload=False)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=omega)
    dm.begin_smoothing(mg)
    h = []
    for c in range(cycles):
        st = dm.smooth(1, mg)
        h.append(st["last_max_update"])
    dm.close()
    return (h[-1] / h[-5]) ** 0.25, h
n = int(sys.argv[1])
for name, mk in [("1x8 plate", lambda: synthetic.cascade(1, 8, 2 * n - 1, n, length=0.125, ay=0.015 / 8)), ("2x8", lambda: synthetic.cascade(2, 8, 2 * n - 1, n, length=0.25, ay=0.015 / 4)),
                 ("8x8", lambda: synthetic.cascade(8, 8, (n + 1) // 2, (n + 3) // 4))]:
    r, h = rate(mk(), 3, 0.8)
    print(f"coarsest={os.environ.get('TM_MG_COARSEST_SWEEPS')} n={n:5d} {name:12s}: factor {r:.3f}  last {h[-1]:.2e}", flush=True)
