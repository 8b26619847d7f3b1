"""GPU vs wall time per multigrid cycle on a 64-block mesh (launch-count sensitivity)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
spec = synthetic.cascade(8, 8, 1025, 513)
dm = smoothing.DeviceMesh(spec, upload=False)
for k, b in enumerate(spec.blocks):
    dm.tfi_block(k, *b.edge_args())
mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
dm.begin_smoothing(mg)
dm.smooth(2, mg)
l0 = smoothing.kernel_launch_count()
t0 = time.perf_counter(); st = dm.smooth(10, mg); t1 = time.perf_counter()
print(f"AA={os.environ.get('TM_MG_AA')} smooth(10): gpu {st['gpu_seconds']*100:.2f} ms/cycle, wall {(t1-t0)*100:.2f} ms/cycle, launches/cycle {(smoothing.kernel_launch_count()-l0)/10}")
t0 = time.perf_counter()
g = 0.0
for _ in range(10):
    g += dm.smooth(1, mg)["gpu_seconds"]
t1 = time.perf_counter()
print(f"10 x smooth(1): gpu {g*100:.2f} ms/cycle, wall {(t1-t0)*100:.2f} ms/cycle")
dm.close()
