import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Mesh

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

os.environ["TM_MG_NESTED"] = "1"
fixed = lambda s: Mesh(blocks=s.blocks, names=s.names, connections=[], boundary_conditions=[])
for n in (65, 129, 257, 513, 1025):
    for name, mk in [("1x1 fixed", lambda: fixed(synthetic.cascade(1, 1, 2 * n - 1, n))), ("1x1 sliding", lambda: synthetic.cascade(1, 1, 2 * n - 1, n)),
                     ("2x2", lambda: synthetic.cascade(2, 2, n, (n + 1) // 2)), ("1x8 plate", lambda: synthetic.cascade(1, 8, 2 * n - 1, n, length=0.125))]:
        r, h = rate(mk(), 3, 0.8)
        print(f"n={n:5d} {name:12s}: factor {r:.3f}  last {h[-1]:.2e}", flush=True)
