"""FAS multigrid on one O4H passage (config 4 as named) at a given cell-count factor: per-cycle convergence (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import load_fixture
from inputgen import passages
from turbomesh_b200 import smoothing
factor = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ds = float(sys.argv[2]) if len(sys.argv) > 2 else None
spec0, z, meta = load_fixture("t106_white")
up, down = z["b0_x_i_min"].copy(), z["b1_x_i_min"].copy()
x0 = min(up[:, 0].min(), down[:, 0].min()); up[:, 0] -= x0; down[:, 0] -= x0
from inputgen.templates import NumCells
o = int(os.environ.get("OGRID", "40"))
th = float(os.environ.get("THICK", "0.001"))
nc = NumCells(o_grid=o, middle_i=100, in_up_j=30, in_down_j=10, in_i=10, out_up_j=40, out_down_j=10, out_i=10, down_j=40, bulge=40, upstream_i=20, downstream_i=10)
mesh, owner = passages.o4h_passages(up, down, meta["pitch"], n_passages=1, factor=factor, o_grid_delta_s=ds, num_cells=nc, o_grid_thickness=th)
plan = smoothing.mg_plan(mesh)
print("factor", factor, "nodes", sum(a * b for a, b in plan[0]), "levels", len(plan), [sum(a * b for a, b in lv) for lv in plan], flush=True)
with smoothing.DeviceMesh(mesh, upload=False) as dm:
    for k, b in enumerate(mesh.blocks):
        dm.tfi_block(k, *b.edge_args())
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=int(os.environ.get("NU", "3")), omega=float(os.environ.get("OMEGA", "0.8")))
    dm.begin_smoothing(mg)
    hist = []
    for c in range(int(os.environ.get("CYCLES", "30"))):
        st = dm.smooth(1, mg)
        fin = bool(np.isfinite(dm.download_block(0)).all())
        hist.append((st["last_max_update"], fin))
        if not fin:
            break
    print(" ".join(f"{u:.1e}{'' if f else '!NaN'}" for u, f in hist), flush=True)
    probe = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=1.0)
    print("jacobi update of the result:", dm.smooth(1, probe)["last_max_update"])
