"""A few sweeps on one O4H passage (config 4 as named, one GPU's share) -- the command the ncu captures of the sweep kernel run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import load_fixture
from inputgen import passages
from turbomesh_b200 import smoothing
factor = int(sys.argv[1]) if len(sys.argv) > 1 else 48
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec0, z, meta = load_fixture("t106_white")
up, down = z["b0_x_i_min"].copy(), z["b1_x_i_min"].copy()
x0 = min(up[:, 0].min(), down[:, 0].min()); up[:, 0] -= x0; down[:, 0] -= x0
mesh, owner = passages.o4h_passages(up, down, meta["pitch"], n_passages=1, factor=factor)
with smoothing.DeviceMesh(mesh, upload=False) as dm:
    for k, b in enumerate(mesh.blocks):
        dm.tfi_block(k, *b.edge_args())
    sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=0.9)
    dm.begin_smoothing(sol)
    for _ in range(2):
        st = dm.smooth(1, sol)
    n = st["nodes"]
    per = st["gpu_seconds"] / sweeps
    print(f"passage factor {factor}: {n} nodes, sweep {per*1e3:.3f} ms = {n*32/per/1e9:.0f} GB/s algorithmic (32 B/node-update)")
