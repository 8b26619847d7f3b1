"""Why does the FAS multigrid stall on an O4H passage?  (diagnostic: relax only, two-grid, deep hierarchy, Picard warm-up)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import load_fixture
from inputgen import passages
from turbomesh_b200 import smoothing
factor = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "mg"
spec0, z, meta = load_fixture("t106_white")
up, down = z["b0_x_i_min"].copy(), z["b1_x_i_min"].copy()
x0 = min(up[:, 0].min(), down[:, 0].min()); up[:, 0] -= x0; down[:, 0] -= x0
mesh, owner = passages.o4h_passages(up, down, meta["pitch"], n_passages=1, factor=factor)
with smoothing.DeviceMesh(mesh, upload=False) as dm:
    for k, b in enumerate(mesh.blocks):
        dm.tfi_block(k, *b.edge_args())
    probe = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=1.0)
    def jac():
        return dm.smooth(1, probe)["last_max_update"]
    if mode == "relax":
        sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=200, omega=0.9)
        dm.begin_smoothing(sol)
        print("relax x200:", " ".join(f"{dm.smooth(1, sol)['last_max_update']:.1e}" for _ in range(15)))
    else:
        warm = int(os.environ.get("PICARD", "0"))
        if warm:
            pic = smoothing.CudaSolver(method="picard_bicgstab", rtol=1e-4, atol=1e-10, max_inner_iterations=500)
            dm.begin_smoothing(pic)
            print("picard:", " ".join(f"{dm.smooth(1, pic)['last_max_update']:.1e}" for _ in range(warm)))
        mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=int(os.environ.get("NU", "3")), omega=0.8)
        if not warm:
            dm.begin_smoothing(mg)
        h = []
        for c in range(25):
            h.append(dm.smooth(1, mg)["last_max_update"])
        print(f"mg levels<={os.environ.get('TM_MG_MAX_LEVELS','all')} coarsest={os.environ.get('TM_MG_COARSEST_SWEEPS','auto')} picard={warm}:", " ".join(f"{v:.1e}" for v in h), "| jacobi", f"{jac():.1e}")
