"""Extended-precision truth of the exact Picard sequence for the parity fixtures -- run anywhere (needs no reference checkout):

    python tests/golden/make_truth.py [t106_white ls89x4_white]

`tests/independent.py` (plain numpy, written from SURVEY.md Appendix A) evaluates smooth.zig:104-154 in x87 80-bit
arithmetic (np.longdouble, 64-bit mantissa): assembly and the White update in that precision, each linear system by a
sparse LU of its fp64 rounding refined with longdouble residuals until the correction is < 1e-19.  The result, rounded
to fp64, is stored per fixture as `<name>_truth.npz`:
  truth{n}_b{k}   mesh after n outer iterations (n in `snapshots`)
  meta            JSON: per outer iteration the distance of the same evaluation in plain fp64 (direct LU, no Krylov
                  tolerance involved) from the truth -- the rounding floor of fp64 on this configuration --, the distance of
                  the truth from itself when every interior input coordinate is moved by +-1 ulp (the conditioning of the
                  exact problem), the smallest wall distance the White update produces at the leading-edge node.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import independent as ind  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from util import chord_of, load_fixture  # noqa: E402

from turbomesh_b200 import synthetic  # noqa: E402


def make(name, snapshots):
    spec, z, meta = load_fixture(name)
    mesh = synthetic.materialize(spec, orc.tfi)
    chord = chord_of(mesh)
    white = (meta["ds_target"], meta["theta_target"]) if meta["control_function"] == "white" else None
    rng = np.random.default_rng(1)
    moved = mesh.copy()
    for b in moved.blocks:  # interior nodes only: the copies of interface nodes must stay coincident
        p = b.points[1:-1, 1:-1]
        s = rng.integers(-1, 2, size=p.shape)
        b.points[1:-1, 1:-1] = np.where(s > 0, np.nextafter(p, np.inf), np.where(s < 0, np.nextafter(p, -np.inf), p))
    T = ind.IndependentSystem(mesh, dtype=np.longdouble, white=white)
    M = ind.IndependentSystem(moved, dtype=np.longdouble, white=white)
    D = ind.IndependentSystem(mesh, dtype=np.float64, white=white)
    out, rows = {}, []
    t0 = time.time()
    for it in range(1, meta["iterations"] + 1):
        T.step(); M.step(); D.step()
        row = {"iteration": it, "fp64_direct_vs_truth": float(np.abs(T.xy - D.xy.astype(np.longdouble)).max()),
               "truth_vs_truth_of_inputs_moved_1ulp": float(np.abs(T.xy - M.xy).max()), "refinement_last_correction": max(T.last_refinement)}
        if white is not None:
            c, _, _, jp1, _ = T._le_frame()
            row["leading_edge_first_cell"] = float(np.sqrt(((jp1 - c) ** 2).sum()))
        rows.append(row)
        print(name, row, f"{time.time() - t0:.0f}s", flush=True)
        if it in snapshots:
            for k, b in enumerate(T.blocks()):
                out[f"truth{it}_b{k}"] = b.astype(np.float64)
    out["meta"] = np.array(json.dumps({"name": name, "chord": chord, "snapshots": sorted(snapshots), "per_iteration": rows,
                                       "how": "tests/independent.py in np.longdouble (x87 80-bit), sparse LU + refinement to < 1e-19"}))
    path = os.path.join(HERE, name + "_truth.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["t106_white", "ls89x4_white"]
    if "t106_white" in which:
        make("t106_white", {8, 10})
    if "ls89x4_white" in which:
        make("ls89x4_white", {5, 10})
