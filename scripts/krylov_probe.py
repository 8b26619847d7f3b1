"""Picard/BiCGStab path on the reference's configurations: device time, Krylov iterations, per-iteration time (tuning aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import load_fixture
from turbomesh_b200 import smoothing, synthetic
which = sys.argv[1:] or ["t106", "ls89", "cuts"]
sol = smoothing.CudaSolver(method="picard_bicgstab")
def run(name, dm, nb, cf, its, reps=3):
    best = None
    for rep in range(reps):
        for k in range(nb):
            dm.tfi_block_resident(k)
        dm.begin_smoothing(sol, cf)
        t0 = time.perf_counter(); st = dm.smooth(its, sol, cf); t1 = time.perf_counter()
        if best is None or st["gpu_seconds"] < best["gpu_seconds"]:
            best = st
    it = best["inner_iterations"] / 2
    print(f"{name}: {best['nodes']} nodes, {best['gpu_seconds']*1e3:.1f} ms device, {best['inner_iterations']} Krylov iterations (x+y), {best['operator_applications']} operator applications, "
          f"converged={best['converged']}, {best['gpu_seconds']*1e6/max(it,1):.1f} us per lock-step iteration, {best['nodes']*best['operator_applications']/best['gpu_seconds']:.3e} node-updates/s", flush=True)
for name in ("t106_white", "ls89x4_white"):
    if name[:4] not in which and name[:4].rstrip("_") not in which:
        continue
    spec, z, meta = load_fixture(name)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    with smoothing.DeviceMesh(spec, upload=False) as dm:
        for k, b in enumerate(spec.blocks):
            dm.tfi_block(k, *b.edge_args())
        run(name, dm, len(spec.blocks), cf, meta["iterations"])
if "cuts" in which:
    base, z, meta = load_fixture("t106_white")
    n = int(os.environ.get("CUTS", "128"))
    batch, groups = synthetic.batch_of_cuts(base, [1.0 + 0.2 * k / max(n - 1, 1) for k in range(n)])
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    with smoothing.DeviceMesh(batch, upload=False) as dm:
        for k, b in enumerate(batch.blocks):
            dm.tfi_block(k, *b.edge_args())
        dm.set_white_groups(groups)
        run(f"{n} cuts", dm, len(batch.blocks), cf, meta["iterations"], reps=2)
