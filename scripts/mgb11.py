import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Mesh
def rate(spec, nu=3, omega=0.8, cycles=14):
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
nobc = lambda s: Mesh(blocks=s.blocks, names=s.names, connections=s.connections, boundary_conditions=[])
for name, mk in [("8x8", lambda: synthetic.cascade(8, 8, 129, 65)), ("8x8 fixed inlet/outlet", lambda: nobc(synthetic.cascade(8, 8, 129, 65))),
                 ("8x2", lambda: synthetic.cascade(8, 2, 129, 257)), ("8x1 walls", lambda: synthetic.cascade(8, 1, 129, 513)),
                 ("4x4", lambda: synthetic.cascade(4, 4, 257, 129)), ("4x4 fixed", lambda: nobc(synthetic.cascade(4, 4, 257, 129)))]:
    r, h = rate(mk())
    print(f"coarsest={os.environ.get('TM_MG_COARSEST_SWEEPS')} {name:24s}: factor {r:.3f}  last {h[-1]:.2e}", flush=True)
