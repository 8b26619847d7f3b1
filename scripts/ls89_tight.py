"""LS89 x4 + White, tight tolerances: error against the extended-precision truth for the Krylov paths (diagnostic)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import load_fixture, chord_of, GOLDEN
from turbomesh_b200 import smoothing, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "ls89x4_white"
spec, z, meta = load_fixture(name)
tz = np.load(os.path.join(GOLDEN, name + "_truth.npz"))
cf = smoothing.White(meta["ds_target"], meta["theta_target"])
for atol, pol in ((1e-13, 0), (1e-13, 1), (1e-13, 2), (1e-13, 4)):
    mesh = synthetic.materialize(spec, smoothing.tfi_block)
    with smoothing.DeviceMesh(mesh) as dm:
        sol = smoothing.CudaSolver.tight(atol=atol, inner_refinement_cycles=pol)
        dm.begin_smoothing(sol, cf)
        rec = []
        for it in range(10):
            st = dm.smooth(1, sol, cf)
            try:
                c = dm.component_stats(0)
                rec.append((c["restarts"], sum(c["iterations"]), max(c["norm_r"]), c["status"]))
            except Exception as e:
                rec.append((st["inner_iterations"], st["last_inner_residual"], st["converged"]))
        blocks = [dm.download_block(k) for k in range(len(mesh.blocks))]
    chord = chord_of(mesh)
    err = max(float(np.abs(b - tz[f"truth10_b{k}"]).max()) for k, b in enumerate(blocks))
    print(f"{os.environ.get('TM_KRYLOV','default')} atol {atol:g} polish {pol}: {sum(r[1] for r in rec if len(r)==4)} its,  err vs truth {err/chord:.2e} chord, statuses {[r[-1] for r in rec]}", flush=True)
