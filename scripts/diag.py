import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from turbomesh_b200 import smoothing, synthetic

def md(a, b): return max(float(np.abs(x.points - y.points).max()) for x, y in zip(a.blocks, b.blocks))
for args in [(2, 2, 40, 24), (4, 2, 33, 17), (8, 8, 16, 12)]:
    spec = synthetic.cascade(*args)
    cpu0 = synthetic.materialize(spec, orc.tfi)
    for its in (1, 3, 6):
        gpu = cpu0.copy(); cpu = cpu0.copy(); cpu2 = cpu0.copy()
        st = smoothing.smooth_mesh(gpu, its, smoothing.CudaSolver.tight())
        so = orc.smooth_mesh(cpu, its, orc.tight_options())
        so2 = orc.smooth_mesh(cpu2, its, orc.tight_options(solver="bicgstab", preconditioner="diagonal"))
        print(args, its, "gpu-vs-gmres %.3e  gmres-vs-bicg(cpu) %.3e gpu-vs-bicg(cpu) %.3e" % (md(gpu, cpu), md(cpu, cpu2), md(gpu, cpu2)),
              "| gpu inner", st["inner_iterations"], "conv", st["converged"], "res %.2e" % st["last_inner_residual"], "| cpu kry", so["krylov_iterations"], so["not_converged"], "bicg", so2["krylov_iterations"], so2["not_converged"])
