"""Configs 1-2 (T106, LS89 x4) through the GPU path with the reference's settings (10 outer iterations, White, rtol 1e-6) and
with tight tolerances: device time of the smoothing call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import load_fixture
from turbomesh_b200 import smoothing, synthetic
for name in ("t106_white", "ls89x4_white"):
    spec, z, meta = load_fixture(name)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    for label, sol in (("reference tolerances (rtol 1e-6, atol 1e-8)", smoothing.CudaSolver(method="picard_bicgstab")), ("tight (atol 1e-13)", smoothing.CudaSolver.tight())):
        mesh = synthetic.materialize(spec, smoothing.tfi_block)
        with smoothing.DeviceMesh(mesh) as dm:
            for rep in range(2):
                for k, b in enumerate(spec.blocks):
                    dm.tfi_block(k, *b.edge_args())
                dm.begin_smoothing(sol, cf)
                t0 = time.perf_counter(); st = dm.smooth(meta["iterations"], sol, cf); t1 = time.perf_counter()
        print(f"{name} ({st['nodes']} nodes) {label}: {st['gpu_seconds']:.3f} s device ({t1-t0:.3f} s wall), {st['inner_iterations']} Krylov iterations, {st['operator_applications']} operator applications, converged={st['converged']}")
