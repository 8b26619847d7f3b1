import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
spec = synthetic.cascade(8, 8, 1025, 513)
dm = smoothing.DeviceMesh(spec, upload=False)
for k, b in enumerate(spec.blocks):
    dm.tfi_block(k, *b.edge_args())
mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8)
dm.begin_smoothing(mg)
h, t = [], 0.0
for c in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    st = dm.smooth(1, mg)
    h.append(st["last_max_update"]); t += st["gpu_seconds"]
    if h[-1] < 1e-10:
        break
print(f"AA={os.environ.get('TM_MG_AA')} 8x8 blocks of 1025x513: {len(h)} cycles, {t*1e3:.1f} ms:", " ".join(f"{v:.1e}" for v in h))
dm.close()
