"""Where the time of the persistent Krylov kernel goes (TM_KRYLOV_TIMING: cycle counters per CTA, printed by the library on
stderr): one outer iteration of T106 + White and of LS89 x4 + White at the reference's tolerances."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import load_fixture
from turbomesh_b200 import smoothing, synthetic
for name in ("t106_white", "ls89x4_white"):
    spec, z, meta = load_fixture(name)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver()
    mesh = synthetic.materialize(spec, smoothing.tfi_block)
    with smoothing.DeviceMesh(mesh) as dm:
        dm.begin_smoothing(sol, cf)
        dm.smooth(2, sol, cf)
        os.environ["TM_KRYLOV_TIMING"] = "1"
        st = dm.smooth(1, sol, cf)
        os.environ.pop("TM_KRYLOV_TIMING")
        print(name, "third outer iteration:", f"{st['gpu_seconds']*1e3:.2f} ms,", st["inner_iterations"], "Krylov iterations (x+y),", st["operator_applications"], "applications", flush=True)
