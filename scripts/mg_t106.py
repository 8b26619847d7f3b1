"""Multigrid on the reference T106 O4H topology against the Picard fixed point (development aid; the test lives in tests/test_gpu_o4h.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from util import load_fixture, chord_of
from turbomesh_b200 import smoothing, synthetic
for name in ("t106_laplace",):
    spec, z, meta = load_fixture(name)
    mesh = synthetic.materialize(spec, smoothing.tfi_block)
    with smoothing.DeviceMesh(mesh) as dm:
        mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
        dm.begin_smoothing(mg)
        h = []
        for c in range(400):
            st = dm.smooth(1, mg)
            h.append(st["last_max_update"])
            if h[-1] < 1e-13: break
        print(name, "blocks", [b.points.shape[:2] for b in mesh.blocks])
        print(f"  multigrid: {len(h)} cycles, ops {st['operator_applications']}/cycle:", " ".join(f"{v:.1e}" for v in h[:12]), "...", " ".join(f"{v:.1e}" for v in h[-4:]))
        mgm = [dm.download_block(k) for k in range(len(mesh.blocks))]
    # Picard to convergence for comparison (Laplace)
    mesh2 = synthetic.materialize(spec, smoothing.tfi_block)
    st = smoothing.smooth_mesh(mesh2, 60, smoothing.CudaSolver.tight(stop_max_update=1e-13))
    err = max(float(np.abs(a - b.points).max()) for a, b in zip(mgm, mesh2.blocks))
    print(f"  picard: {st['outer_iterations']} outer its, last update {st['last_max_update']:.1e}; |mg - picard| = {err:.2e} (chord {chord_of(mesh2):.3g})")
