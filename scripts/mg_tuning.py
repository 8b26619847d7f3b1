"""Cycle counts of the multi-block multigrid on the N-column tiling (config 4 shape at reduced block size, one GPU): which knob
removes the 13 -> 28 cycle growth between N <= 2 and N >= 4 (plate tips inside the passage)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
ni, nj = int(os.environ.get("BNI", "1025")), int(os.environ.get("BNJ", "513"))
for world in (2, 4, 8):
    for nu, omega in ((3, 0.8), (4, 0.8), (3, 0.9)):
        spec = synthetic.cascade(world, 8, ni, nj, length=world / 8.0, ay=0.015 * world / 8.0)
        with smoothing.DeviceMesh(spec, upload=False) as dm:
            for k, b in enumerate(spec.blocks):
                dm.tfi_block(k, *b.edge_args())
            mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=omega, stop_max_update=1e-10)
            dm.begin_smoothing(mg)
            st = dm.smooth(100, mg)
        print(f"AA={os.environ.get('TM_MG_AA', '1')} window={os.environ.get('TM_MG_AA_WINDOW', '3')} coarsest={os.environ.get('TM_MG_COARSEST_SWEEPS', 'auto')} "
              f"N={world} nu={nu} omega={omega}: {st['outer_iterations']} cycles, {st['gpu_seconds']*1e3:.1f} ms, ops {st['operator_applications']}, last {st['last_max_update']:.1e}", flush=True)
