"""A few cycles of the single-block (non-nested) multigrid on 8192^2 -- the command for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
spec = synthetic.single_block(8192, 8192)
dm = smoothing.DeviceMesh(spec, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
dm.begin_smoothing(mg)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
st = dm.smooth(n, mg)
print(f"{n} cycles: {st['gpu_seconds']*1e3/n:.3f} ms/cycle")
dm.close()
