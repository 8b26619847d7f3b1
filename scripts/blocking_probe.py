"""O4H blocking of a batch of cuts on the device (turbomesh_b200/blocking.py): wall time per batch size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import load_fixture
from turbomesh_b200.blocking import Cells, Cut, O4HBatch
from turbomesh_b200.clustering import Roberts
spec, z, meta = load_fixture("t106_white")
cells = Cells(o_grid=40, middle_i=100, in_up_j=30, in_down_j=10, in_i=10, out_up_j=40, out_down_j=10, out_i=10, down_j=40, bulge=40, upstream_i=20, downstream_i=10)
b = O4HBatch(cells, Roberts(0.5, 1.03))
for n in (1, 16, 128, 1024):
    cuts = [Cut(z["b0_x_i_min"] * s, z["b1_x_i_min"] * s, float(meta["pitch"]) * s) for s in [1.0 + 0.2 * k / max(n - 1, 1) for k in range(n)]]
    b.run(cuts[:1])
    t0 = time.perf_counter(); mesh, _ = b.run(cuts); dt = time.perf_counter() - t0
    print(f"{n:5d} cuts: {dt*1e3:8.1f} ms ({dt/n*1e6:7.1f} us per cut), {len(mesh.blocks)} blocks, {mesh.num_nodes()} nodes", flush=True)
