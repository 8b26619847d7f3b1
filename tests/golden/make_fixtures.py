"""Generates the golden fixtures of tests/golden/ -- run HERE (the container that mounts /root/reference):

    python tests/golden/make_fixtures.py

Inputs come from the reference's shipped examples (examples/T106/T106.json + T106_ps.dat / T106_ss.dat,
examples/LS89/LS89.json); the O4H template (host restatement in tests/inputgen/templates.py) turns them into block edges,
the CPU oracle (oracle/turbomesh_oracle.c) produces TFI and smoothed meshes.  The reference itself cannot run here
(Zig is not installed), so these are ORACLE outputs: parity stays "unpinned" with respect to the Zig binary.

Fixture layout (npz): for block k the eight TFI inputs `b{k}_x_i_min ... b{k}_t2`; `conn` (n,11) rows
[block0, side0, start0, end0, block1, side1, start1, end1, has_per, px, py]; `bcs` (n,5) rows [block, side, start, end, kind];
`tfi_b{k}` oracle TFI output; `smooth_b{k}` mesh after `iterations` outer iterations of the oracle in tight mode;
`meta` JSON incl. `oracle_spread` = max |dx| between two tight oracle variants (GMRES restart 30 vs 60).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc  # noqa: E402
from turbomesh_b200.discrete import Block2d, Edge  # noqa: E402
from inputgen.input import Input  # noqa: E402

REF = "/root/reference"


class Recorder:
    """Wraps the oracle TFI and records the edge inputs of every Block2d.init call (discrete.zig:142-159)."""

    def __init__(self):
        self.calls = []

    def __call__(self, *args):
        self.calls.append([np.array(a, dtype=np.float64, copy=True) for a in args])
        return orc.tfi(*args)


def topo_arrays(mesh):
    conn = np.array([[c.ranges[0].block, int(c.ranges[0].side), c.ranges[0].start, c.ranges[0].end,
                      c.ranges[1].block, int(c.ranges[1].side), c.ranges[1].start, c.ranges[1].end,
                      0 if c.periodicity is None else 1, 0.0 if c.periodicity is None else c.periodicity[0],
                      0.0 if c.periodicity is None else c.periodicity[1]] for c in mesh.connections], dtype=np.float64)
    bcs = np.array([[b.range.block, int(b.range.side), b.range.start, b.range.end, int(b.kind)] for b in mesh.boundary_conditions], dtype=np.int64)
    return conn, bcs


def make(name, json_path, cell_factor, iterations, control_function, store_tfi=True):
    inp = Input.from_json(open(json_path).read())
    if cell_factor != 1:
        inp.template.num_cells = inp.template.num_cells.scaled(cell_factor)
    geom = inp.geometry(REF)
    rec = Recorder()
    mesh = inp.template.run(geom, tfi=rec)
    out = {}
    names = ("x_i_min", "x_i_max", "x_j_min", "x_j_max", "s1", "s2", "t1", "t2")
    for k, call in enumerate(rec.calls):
        for n, a in zip(names, call):
            out[f"b{k}_{n}"] = a
        if store_tfi:
            out[f"tfi_b{k}"] = mesh.blocks[k].points.copy()
    out["conn"], out["bcs"] = topo_arrays(mesh)
    wcf = inp.smoothing.wall_control_function
    (tag, val), = wcf.items()
    kw = dict(control_function=control_function)
    if control_function == "white":
        kw.update(ds_target=float(val["ds_target"]), theta_target=float(val.get("theta_target", 0.5 * np.pi)))
    t0 = time.time()
    variant = mesh.copy()
    st = orc.smooth_mesh(mesh, iterations, orc.tight_options(max_iters=100000, **kw))
    print(f"{name}: {mesh.num_nodes()} nodes, oracle tight {iterations} its in {time.time() - t0:.1f}s, krylov {st['krylov_iterations']}, not converged {st['not_converged']}")
    assert st["not_converged"] == 0
    # Self-consistency of the oracle: the same tight solve with GMRES restart 60 instead of 30.  Both "converge" to
    # 1e-15, yet the meshes differ by kappa * eps * |x| -- the fp64 floor below which no parity claim is meaningful.
    st2 = orc.smooth_mesh(variant, iterations, orc.tight_options(max_iters=100000, restart=60, **kw))
    assert st2["not_converged"] == 0
    spread = max(float(np.abs(a.points - b.points).max()) for a, b in zip(mesh.blocks, variant.blocks))
    print(f"  oracle spread (restart 30 vs 60): {spread:.3e}")
    for k, b in enumerate(mesh.blocks):
        out[f"smooth_b{k}"] = b.points
    out["meta"] = np.array(json.dumps({"name": name, "iterations": iterations, "control_function": control_function,
                                       "ds_target": kw.get("ds_target", 0.0), "theta_target": kw.get("theta_target", 0.0),
                                       "pitch": geom.pitch, "nodes": mesh.num_nodes(), "cell_factor": cell_factor,
                                       "oracle": "gmres+ilu0 rtol=atol=1e-15", "oracle_spread": spread}))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("  wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["t106_white", "t106_laplace", "ls89x4_white"]
    if "t106_white" in which:
        make("t106_white", f"{REF}/examples/T106/T106.json", 1, 10, "white")
    if "t106_laplace" in which:
        make("t106_laplace", f"{REF}/examples/T106/T106.json", 1, 4, "laplace", store_tfi=False)
    if "ls89x4_white" in which:
        make("ls89x4_white", f"{REF}/examples/LS89/LS89.json", 2, 10, "white", store_tfi=False)
