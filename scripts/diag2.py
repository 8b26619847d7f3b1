import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, time
from turbomesh_b200 import smoothing, synthetic
from util import load_fixture, chord_of
name = sys.argv[1]; atol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-13
spec, z, meta = load_fixture(name)
mesh = synthetic.materialize(spec, smoothing.tfi_block)
cf = smoothing.White(meta["ds_target"], meta["theta_target"]) if meta["control_function"] == "white" else smoothing.Laplace()
with smoothing.DeviceMesh(mesh) as dm:
    sol = smoothing.CudaSolver.tight(atol=atol)
    dm.begin_smoothing(sol, cf)
    for it in range(meta["iterations"]):
        t = time.time(); st = dm.smooth(1, sol, cf); dt = time.time() - t
        print(it, "inner", st["inner_iterations"], "ops", st["operator_applications"], "conv", st["converged"], "res %.3e" % st["last_inner_residual"], "upd %.3e" % st["last_max_update"], "%.2fs" % dt)
    dm.download()
err = max(float(np.abs(b.points - z[f"smooth_b{k}"]).max()) for k, b in enumerate(mesh.blocks))
print("err %.3e chord %.4f err/chord %.3e spread %.3e" % (err, chord_of(mesh), err / chord_of(mesh), meta["oracle_spread"]))
