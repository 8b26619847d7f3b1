"""Device time per multigrid cycle on the 1x8-block cascade (67 M nodes); also the command the ncu launch list in profiles/ was taken from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 3
spec = synthetic.cascade(1, 8, 4097, 2049, length=0.125)
dm = smoothing.DeviceMesh(spec, upload=False)
for k, b in enumerate(spec.blocks):
    dm.tfi_block(k, *b.edge_args())
mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
dm.begin_smoothing(mg)
st = dm.smooth(1, mg)
st = dm.smooth(cycles, mg)
print(f"{cycles} cycles: {st['gpu_seconds']*1e3/cycles:.3f} ms/cycle, last_max_update {st['last_max_update']:.3e}, launches {smoothing.kernel_launch_count()}")
dm.close()
