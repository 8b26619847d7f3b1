"""Which topology feature of the O4H passage breaks the multigrid's two-grid convergence?  Variants of one passage with
features switched off (their nodes become fixed walls), and a refined T-junction mesh."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import load_fixture
from inputgen import passages
from turbomesh_b200 import smoothing, synthetic
from turbomesh_b200.discrete import Mesh
factor = int(sys.argv[1]) if len(sys.argv) > 1 else 4
spec0, z, meta = load_fixture("t106_white")
up, down = z["b0_x_i_min"].copy(), z["b1_x_i_min"].copy()
x0 = min(up[:, 0].min(), down[:, 0].min()); up[:, 0] -= x0; down[:, 0] -= x0
base, owner = passages.o4h_passages(up, down, meta["pitch"], n_passages=1, factor=factor)

def run(name, mesh, cycles=16):
    with smoothing.DeviceMesh(mesh, upload=False) as dm:
        for k, b in enumerate(mesh.blocks):
            dm.tfi_block(k, *b.edge_args())
        mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
        dm.begin_smoothing(mg)
        h = [dm.smooth(1, mg)["last_max_update"] for _ in range(cycles)]
        print(f"{name:42s}", " ".join(f"{v:.1e}" for v in h[1:]), f"| rate {(h[-1] / h[-5]) ** 0.25:.2f}", flush=True)

def variant(drop_bcs=False, drop_periodic=False, keep_conns=None):
    m = Mesh(blocks=list(base.blocks), names=list(base.names), connections=[c for k, c in enumerate(base.connections)
             if not (drop_periodic and c.periodicity is not None) and (keep_conns is None or k in keep_conns)],
             boundary_conditions=[] if drop_bcs else list(base.boundary_conditions))
    return m

nu = int(os.environ.get("NU", "3")); om = float(os.environ.get("OMEGA", "0.8"))
def run2(name, mesh, cycles=14):
    with smoothing.DeviceMesh(mesh, upload=False) as dm:
        for k, b in enumerate(mesh.blocks):
            dm.tfi_block(k, *b.edge_args())
        print("   plan:", [lv[:2] for lv in smoothing.mg_plan(mesh)][:12])
        mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=om)
        dm.begin_smoothing(mg)
        h = [dm.smooth(1, mg)["last_max_update"] for _ in range(cycles)]
        print(f"{name:30s} nu={nu} omega={om} levels<={os.environ.get('TM_MG_MAX_LEVELS','all')} coarsest={os.environ.get('TM_MG_COARSEST_SWEEPS','auto')}:", " ".join(f"{v:.1e}" for v in h[1:]), f"| rate {(h[-1] / h[-5]) ** 0.25:.2f}", flush=True)
which = os.environ.get("WHICH", "ogrid")
if which == "ogrid":
    run2("O-grid only", variant(drop_bcs=True, keep_conns={0, 1}))
else:
    run2("outer blocks only", variant(drop_bcs=True, keep_conns=set(range(2, 12))))
