"""Multi-block FAS multigrid experiments: convergence history and fixed-point parity against the relaxation solver."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from turbomesh_b200 import smoothing, synthetic

def run(n_bi, n_bj, ni, nj, owner=None, n_ranks=1, cycles=30, nu=3, omega=0.8, check=True, length=1.0):
    spec = synthetic.cascade(n_bi, n_bj, ni, nj, length=length)
    kw = {} if owner is None else dict(owner=owner, rank=None, n_ranks=n_ranks)
    dm = smoothing.DeviceMesh(spec, upload=False, **kw)
    for k, b in enumerate(spec.blocks):
        dm.tfi_block(k, *b.edge_args())
    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=nu, omega=omega)
    dm.begin_smoothing(mg)
    hist, t = [], 0.0
    for c in range(cycles):
        st = dm.smooth(1, mg)
        hist.append(st["last_max_update"]); t += st["gpu_seconds"]
        if st["last_max_update"] < 1e-12:
            break
    print(f"cascade {n_bi}x{n_bj} blocks of {ni}x{nj} ranks={n_ranks}: {len(hist)} cycles, {t*1e3:.2f} ms, ops/cycle={st['operator_applications']}")
    print("  max_update:", " ".join(f"{h:.2e}" for h in hist))
    if check:
        out = [dm.download_block(k) for k in range(len(spec.blocks))]
        # fixed point check: a long relaxation run from the multigrid result must not move the mesh
        rl = smoothing.CudaSolver(method="relax", sweeps_per_iteration=200, omega=0.9)
        st = dm.smooth(1, rl)
        out2 = [dm.download_block(k) for k in range(len(spec.blocks))]
        print(f"  after 200 more sweeps: moved by {max(float(np.abs(a-b).max()) for a, b in zip(out, out2)):.3e}, last sweep update {st['last_max_update']:.3e}")
    dm.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(1, 1, 65, 33)
        run(2, 1, 65, 33)
        run(2, 2, 65, 33)
        run(2, 2, 65, 33, owner=[0, 0, 1, 1], n_ranks=2)
        run(4, 4, 129, 65)
        run(4, 4, 129, 65, owner=[k // 4 for k in range(16)], n_ranks=4)
    elif which == "big":
        run(1, 8, 4097, 2049, cycles=25, check=False, length=0.125)
        run(1, 8, 4097, 2049, cycles=25, check=False, length=1.0)
