"""The oracle against a second, independently written evaluation of SURVEY.md Appendix A.3-A.9 (tests/independent.py),
and both against hand-computed rows.  CPU only."""
import json
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import independent as ind
from util import GOLDEN, chord_of, load_fixture

from turbomesh_b200 import synthetic
from turbomesh_b200.boundary import Condition, ConditionTag, Connection, Range, Side
from turbomesh_b200.discrete import Mesh


class _Block:
    def __init__(self, points):
        self.points = np.ascontiguousarray(points, dtype=np.float64)


def _mapped(u, v):
    """a smooth, non-affine map of the parameter plane: every metric term (g11, g22, g12) is non-trivial"""
    U, V = np.meshgrid(u, v, indexing="ij")
    return np.stack([U + 0.05 * np.sin(np.pi * U) * np.sin(np.pi * V) + 0.1 * V * V, V + 0.04 * np.sin(2 * np.pi * U) * np.cos(np.pi * V) + 0.07 * U], axis=-1)


def t_junction_mesh():
    """Three blocks meeting in a T: A spans the top, B and C share the bottom; the node (1, 1) is an edge node of A and a
    corner of B and C (junction stencil with 2 + 1 + 1 diagonal neighbours)."""
    m = Mesh()
    k = np.arange(11) / 5.0
    m.blocks = [_Block(_mapped(k, 1.0 + np.arange(6) / 5.0)),            # A: 11 x 6, x in [0, 2], y in [1, 2]
                _Block(_mapped(k[:6], np.arange(6) / 5.0)),              # B:  6 x 6, x in [0, 1], y in [0, 1]
                _Block(_mapped(k[5:], np.arange(6) / 5.0))]              # C:  6 x 6, x in [1, 2], y in [0, 1]
    m.names = ["A", "B", "C"]
    m.connections = [Connection((Range(0, Side.i_min, 0, 5), Range(1, Side.i_max, 0, 5))),
                     Connection((Range(0, Side.i_min, 5, 10), Range(2, Side.i_max, 0, 5))),
                     Connection((Range(1, Side.j_max, 0, 5), Range(2, Side.j_min, 0, 5)))]
    m.boundary_conditions = []
    return m


def _compare(orc, mesh, white=None):
    kw = dict(control_function="white", ds_target=white[0], theta_target=white[1]) if white else {}
    S = ind.IndependentSystem(mesh, white=white)
    O = orc.System(mesh, orc.options(**kw))
    assert np.array_equal(O.kinds(), S.kinds_flat()), "node kinds differ"
    assert len(O.junctions()) == len(S.junctions)
    for a, b in zip(O.junctions(), S.junctions):
        assert a["ids"] == [g for g, _ in b["copies"]]
    O.fill(0)
    assert np.array_equal(O.control_function(), S.cf)
    worst = 0.0
    for y_mode in (False, True):
        O.fill_specific(y_mode)
        p, i, v, rx, ry = O.csr()
        Ao = sp.csr_matrix((v, i, p), shape=(S.n, S.n))
        A, rhs = S.csr(y_mode)
        D = abs(A - Ao)
        scale = abs(Ao).max()
        worst = max(worst, (D.max() if D.nnz else 0.0) / scale)
        assert np.allclose(rhs, ry if y_mode else rx, rtol=0, atol=4e-16 * max(1.0, np.abs(rx).max()))
    assert worst <= 4e-16, f"CSR values differ by {worst:.2e} (relative)"
    O.close()
    return S


def test_hand_computed_interior_rows():
    """A.5 on a 3 x 3 block whose nine coefficients can be written down by hand."""
    # x = 2 i, y = 3 j: x_xi = 1 * 2 / ... central differences: x_xi = (4 - 0) / 2 = 2, y_eta = 3, g11 = 4, g22 = 9, g12 = 0
    i, j = np.meshgrid(np.arange(3.0), np.arange(3.0), indexing="ij")
    m = Mesh(); m.blocks = [_Block(np.stack([2 * i, 3 * j], axis=-1))]; m.connections = []; m.boundary_conditions = []
    A, rhs = ind.IndependentSystem(m).csr(False)
    row = A[4].toarray().ravel()
    assert np.array_equal(row, [0, 9, 0, 4, -26, 4, 0, 9, 0])      # [l-nj-1 .. l+nj+1]: W and E carry g22, S and N carry g11
    # sheared: x = 2 i + j, y = 3 j  ->  x_eta = 1, g11 = 4, g22 = 1 + 9 = 10, g12 = 2; corners -+ g12 / 2
    m.blocks = [_Block(np.stack([2 * i + j, 3 * j], axis=-1))]
    S = ind.IndependentSystem(m)
    S.cf[4] = (0.5, -0.25)                                             # P, Q: E/W = g22 (1 +- P/2), N/S = g11 (1 +- Q/2)
    row = S.csr(False)[0][4].toarray().ravel()
    assert np.allclose(row, [-1.0, 10 * 0.75, 1.0, 4 * 1.125, -28, 4 * 0.875, 1.0, 10 * 1.25, -1.0], rtol=0, atol=1e-15)


def test_oracle_rows_equal_the_independent_evaluation_t_junction(orc):
    S = _compare(orc, t_junction_mesh())
    (jn,) = S.junctions
    assert len(jn["copies"]) == 3
    A, _ = S.csr(False)
    row = A[jn["primary"]]
    assert row.nnz == 5 and row[0, jn["primary"]] == -4.0              # 2 (edge node of A) + 1 + 1 diagonal neighbours


@pytest.mark.parametrize("shape", [(2, 2, 9, 8), (3, 2, 8, 11), (2, 3, 12, 7)])
def test_oracle_rows_equal_the_independent_evaluation_cascade(orc, shape):
    """interface, periodic interface, 4-block corner junctions (periodic ones included), sliding inlet / outlet rows"""
    mesh = synthetic.materialize(synthetic.cascade(*shape), orc.tfi)
    S = _compare(orc, mesh)
    assert S.junctions and any(c.periodicity for c in mesh.connections)


def test_oracle_rows_equal_the_independent_evaluation_t106(orc):
    """the reference's own topology: sub-range and reversed connections, 3- and 5-copy... junctions, T-junctions on the O-grid
    line, periodic pairs, the White control function incl. the leading-edge override"""
    spec, z, meta = load_fixture("t106_white")
    mesh = synthetic.materialize(spec, orc.tfi)
    S = _compare(orc, mesh, white=(meta["ds_target"], meta["theta_target"]))
    assert S.n == 25118 and len(S.junctions) >= 8


def test_white_update_matches_the_oracle(orc):
    """A.9 after a few exact Picard steps: accumulated (P, Q) of both implementations agree to rounding"""
    spec, z, meta = load_fixture("t106_white")
    mesh = synthetic.materialize(spec, orc.tfi)
    white = (meta["ds_target"], meta["theta_target"])
    S = ind.IndependentSystem(mesh, white=white)
    O = orc.System(mesh, orc.tight_options(control_function="white", ds_target=white[0], theta_target=white[1]))
    for _ in range(3):
        S.step()
    for n in range(3):
        O.iterate(n)
    O.fill(3)      # runs White.update on the mesh after three iterations (smooth.zig:1107-1110)
    S.white_update()
    scale = np.abs(S.cf).max()
    assert np.abs(O.control_function() - S.cf).max() <= 1e-9 * scale
    O.close()


def _truth(name):
    z = np.load(os.path.join(GOLDEN, name + "_truth.npz"))
    return z, json.loads(str(z["meta"]))


def test_exact_picard_sequence_against_the_extended_precision_truth(orc):
    """Config 1 (T106 + White).  While the mesh is regular (outer iterations <= 8) plain fp64 reproduces the 80-bit truth to
    well below 1e-9 chord.  From iteration 9 on it cannot: the reference's leading-edge update (wall_control_function.zig:
    394-472, the sign-flipped xi derivative) collapses the first cell of connection 0 to ~2e-13 m, eta = x[1] - x[0] then
    carries 4 digits in fp64 and (P, Q) of that line differ by 6e-5 between fp64 and exact arithmetic."""
    spec, z, meta = load_fixture("t106_white")
    tz, tmeta = _truth("t106_white")
    mesh = synthetic.materialize(spec, orc.tfi)
    chord = chord_of(mesh)
    rows = tmeta["per_iteration"]
    assert max(r["fp64_direct_vs_truth"] for r in rows[:8]) <= 2e-10 * chord
    assert min(r["leading_edge_first_cell"] for r in rows) < 1e-12          # the degenerate cell
    assert rows[9]["fp64_direct_vs_truth"] > 1e-9 * chord                    # the fp64 floor of the full 10 iterations
    assert max(r["truth_vs_truth_of_inputs_moved_1ulp"] for r in rows) <= 1e-9 * chord   # the exact problem is well conditioned
    S = ind.IndependentSystem(mesh, white=(meta["ds_target"], meta["theta_target"]))
    for _ in range(8):
        S.step()
    err8 = max(float(np.abs(b - tz[f"truth8_b{k}"]).max()) for k, b in enumerate(S.blocks()))
    assert err8 <= 1e-9 * chord
    # the oracle (gmres + ilu0, rtol = atol = 1e-15) after 8 and 10 iterations
    om = mesh.copy()
    opts = orc.tight_options(control_function="white", ds_target=meta["ds_target"], theta_target=meta["theta_target"], max_iters=100000)
    orc.smooth_mesh(om, 8, opts)
    err8o = max(float(np.abs(b.points - tz[f"truth8_b{k}"]).max()) for k, b in enumerate(om.blocks))
    assert err8o <= 1e-9 * chord, f"oracle vs truth after 8 iterations: {err8o / chord:.2e} chord"
    err10o = max(float(np.abs(z[f"smooth_b{k}"] - tz[f"truth10_b{k}"]).max()) for k in range(len(mesh.blocks)))
    floor = rows[9]["fp64_direct_vs_truth"]
    print(f"oracle vs truth: {err8o / chord:.2e} chord after 8, {err10o / chord:.2e} chord after 10 iterations (fp64 floor {floor / chord:.2e} chord)")
    assert err10o <= 3.0 * floor
