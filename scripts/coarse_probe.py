"""Two-level preconditioner of the persistent Krylov kernel (krylov_coarse.cuh) on the reference's configurations: device
time, Krylov iterations and the error against the extended-precision truth, coarse space off / on.
Usage: python scripts/coarse_probe.py [t106 ls89 cuts]"""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from util import load_fixture, chord_of, GOLDEN
from turbomesh_b200 import smoothing, synthetic
which = sys.argv[1:] or ["t106", "ls89"]

def run(name, label, env, tight):
    for k, v in env.items():
        os.environ[k] = v
    spec, z, meta = load_fixture(name)
    tz = np.load(os.path.join(GOLDEN, name + "_truth.npz"))
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver.tight() if tight else smoothing.CudaSolver()
    best = None
    for rep in range(3):
        mesh = synthetic.materialize(spec, smoothing.tfi_block)
        with smoothing.DeviceMesh(mesh) as dm:
            dm.begin_smoothing(sol, cf)
            st = dm.smooth(10, sol, cf)
            blocks = [dm.download_block(k) for k in range(len(mesh.blocks))]
        if best is None or st["gpu_seconds"] < best["gpu_seconds"]:
            best = st
    chord = chord_of(mesh)
    err = max(float(np.abs(b - tz[f"truth10_b{k}"]).max()) for k, b in enumerate(blocks))
    print(f"{name} {label:28s} {'tight' if tight else 'reference tolerances'}: {best['gpu_seconds']*1e3:8.2f} ms, {best['inner_iterations']:6d} iterations (x+y), "
          f"{best['operator_applications']:6d} applications, converged={best['converged']}, residual {best['last_inner_residual']:.2e}, err vs truth {err/chord:.2e} chord", flush=True)
    for k in env:
        os.environ.pop(k)

cases = [("jacobi", {"TM_KRYLOV_COARSE": "0"}), ("coarse (default)", {"TM_KRYLOV_COARSE": "1"}), ("coarse, tables in L2", {"TM_KRYLOV_COARSE": "1", "TM_KRYLOV_COARSE_NOCACHE": "1"}),
         ("coarse 16x16", {"TM_KRYLOV_COARSE": "1", "TM_KRYLOV_COARSE_PATCH": "16x16"}), ("coarse, every 3rd", {"TM_KRYLOV_COARSE": "1", "TM_KRYLOV_COARSE_EVERY": "3"}), ("coarse, once", {"TM_KRYLOV_COARSE": "1", "TM_KRYLOV_COARSE_EVERY": "100"})]
if os.environ.get("CASES"):
    cases = [c for c in cases if c[0] in os.environ["CASES"].split(";")]
for name in ("t106_white", "ls89x4_white"):
    if name[:4] not in which:
        continue
    for tight in (False, True):
        for label, env in cases:
            run(name, label, env, tight)
