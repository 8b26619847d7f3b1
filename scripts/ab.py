"""A/B of the two interior kernels (register-only vs bulk-async ring): bit-identical results, device time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["regs", "bulk"]
res = {}
for spec_name, spec in (("single", synthetic.single_block(n, n)), ("cascade", synthetic.cascade(4, 2, max(n // 8, 40), max(n // 16, 24)))):
    for var in variants:
        name, _, rows = var.partition(":")
        os.environ["TM_INTERIOR"] = name
        os.environ["TM_TILE_ROWS"] = rows or "32"
        dm = smoothing.DeviceMesh(spec, upload=False)
        for k, b in enumerate(spec.blocks):
            dm.tfi_block(k, *b.edge_args())
        sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=0.9)
        dm.begin_smoothing(sol)
        best = 1e9
        for _ in range(4):
            st = dm.smooth(1, sol)
            best = min(best, st["gpu_seconds"] / sweeps)
        nodes = dm.node_count
        out = [dm.download_block(k) for k in range(len(spec.blocks))] if nodes < 5e7 else None
        res[(spec_name, var)] = out
        print(f"{spec_name:8s} {var:10s} nodes {nodes:10d} sweep {best*1e3:8.4f} ms  {nodes/best/1e9:7.1f} G upd/s  {nodes*32/best/1e9:7.1f} GB/s  max_update {st['last_max_update']:.3e}")
        dm.close()
    ref = res[(spec_name, variants[0])]
    for var in variants[1:]:
        o = res[(spec_name, var)]
        if ref is not None and o is not None:
            print(f"   {spec_name} {variants[0]} vs {var}: identical = {all(np.array_equal(a, b) for a, b in zip(ref, o))}, max|diff| = {max(float(np.abs(a - b).max()) for a, b in zip(ref, o)):.3e}")
