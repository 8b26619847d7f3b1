"""Convergence factor per V-cycle of the multi-block multigrid over a set of topologies and smoother settings (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Mesh

def rate(spec, nu, omega, cycles=16):
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
    return (h[-1] / h[-5]) ** 0.25, h[-1]

os.environ["TM_MG_NESTED"] = "1"
for name, mk in [("1x1 fixed", lambda: (lambda s: Mesh(blocks=s.blocks, names=s.names, connections=[], boundary_conditions=[]))(synthetic.cascade(1, 1, 257, 129))),
                 ("1x1 sliding", lambda: synthetic.cascade(1, 1, 257, 129)),
                 ("2x1", lambda: synthetic.cascade(2, 1, 129, 129)),
                 ("1x2 periodic", lambda: synthetic.cascade(1, 2, 257, 65)),
                 ("2x2", lambda: synthetic.cascade(2, 2, 129, 65)),
                 ("4x4", lambda: synthetic.cascade(4, 4, 65, 33))]:
    for nu, om in [(3, 0.8), (2, 0.8), (3, 0.9), (3, 0.7)]:
        r, last = rate(mk(), nu, om)
        print(f"{name:14s} V({nu},{nu}) omega {om}: factor {r:.3f} (last {last:.2e})", flush=True)
