"""Shared helpers of the test-suite: golden fixtures (tests/golden/*.npz) -> meshes."""
import json
import os

import numpy as np

from turbomesh_b200.boundary import Condition, ConditionTag, Connection, Range, Side
from turbomesh_b200.discrete import Edge, Mesh
from turbomesh_b200.synthetic import EdgeBlock

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    """Returns (mesh of EdgeBlocks, npz dict, meta dict) of a fixture written by tests/golden/make_fixtures.py."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    mesh = Mesh()
    k = 0
    while f"b{k}_s1" in z:
        e = {n: z[f"b{k}_{n}"] for n in ("x_i_min", "x_i_max", "x_j_min", "x_j_max", "s1", "s2", "t1", "t2")}
        mesh.add_block(f"b{k}", EdgeBlock(Edge(e["x_i_min"], e["s1"]), Edge(e["x_i_max"], e["s2"]), Edge(e["x_j_min"], e["t1"]), Edge(e["x_j_max"], e["t2"])))
        k += 1
    for row in z["conn"]:
        per = (float(row[9]), float(row[10])) if row[8] else None
        mesh.connections.append(Connection((Range(int(row[0]), Side(int(row[1])), int(row[2]), int(row[3])),
                                            Range(int(row[4]), Side(int(row[5])), int(row[6]), int(row[7]))), per))
    for row in z["bcs"]:
        mesh.boundary_conditions.append(Condition(Range(int(row[0]), Side(int(row[1])), int(row[2]), int(row[3])), ConditionTag(int(row[4]))))
    return mesh, z, meta


def max_diff(a, b):
    return max(float(np.abs(x.points - y.points).max()) for x, y in zip(a.blocks, b.blocks))


def chord_of(mesh):
    """Axial extent of the blade wall (block 0, side i_min) -- the length scale of the parity tolerance."""
    w = mesh.blocks[0].points[:, 0, :] if mesh.blocks[0].points is not None else mesh.blocks[0].i_min.points
    return float(w[:, 0].max() - w[:, 0].min())
