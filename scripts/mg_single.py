import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
spec = synthetic.single_block(8192, 8192)
dm = smoothing.DeviceMesh(spec, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
for nu in (3, 2):
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=0.8)
    for rep in range(2):
        dm.tfi_block_resident(0)
        dm.begin_smoothing(mg)
        h, t = [], 0.0
        for c in range(40):
            st = dm.smooth(1, mg)
            h.append(st["last_max_update"]); t += st["gpu_seconds"]
            if h[-1] < 1e-10: break
    print(f"AA={os.environ.get('TM_MG_AA')} V({nu},{nu}): {len(h)} cycles, {t*1e3:.1f} ms:", " ".join(f"{v:.1e}" for v in h))
dm.close()
