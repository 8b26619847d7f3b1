"""CPU tests that earn the oracle its trust (PARITY UNPINNED against the Zig binary, see oracle/turbomesh_oracle.c):
analytic identities, an independent scipy sparse-LU solve of the assembled system, the reference's adjacent
known-answer vectors and its runtime invariants."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import load_fixture, max_diff  # noqa: E402

from turbomesh_b200 import synthetic  # noqa: E402
from turbomesh_b200.boundary import Condition, ConditionTag, Connection, Range, Side  # noqa: E402
from turbomesh_b200.clustering import Roberts, SingleHyperbolicClustering, Uniform  # noqa: E402
from turbomesh_b200.discrete import Block2d, Mesh  # noqa: E402


# ---------------------------------------------------------------------------------------------- TFI
def _line(a, b, u):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return a[None, :] + u[:, None] * (b - a)[None, :]


def test_tfi_reproduces_bilinear_patch(orc):
    """Straight edges + identical clusterings on opposite sides => TFI is the bilinear map (tfi.zig:185-197)."""
    ni, nj = 17, 11
    s, t = Roberts(0.5, 1.1).compute(ni), SingleHyperbolicClustering(0.02).compute(nj)
    s[0], s[-1], t[0], t[-1] = 0, 1, 0, 1
    c00, cn0, c0m, cnm = (0.1, -0.2), (2.0, 0.3), (-0.4, 1.5), (2.5, 2.2)
    out = orc.tfi(_line(c00, cn0, s), _line(c0m, cnm, s), _line(c00, c0m, t), _line(cn0, cnm, t), s, s, t, t)
    S, T = np.meshgrid(s, t, indexing="ij")
    exact = ((1 - S) * (1 - T))[..., None] * np.array(c00) + (S * (1 - T))[..., None] * np.array(cn0) \
        + ((1 - S) * T)[..., None] * np.array(c0m) + (S * T)[..., None] * np.array(cnm)
    assert np.abs(out - exact).max() < 1e-14


def test_tfi_boundary_nodes_equal_edges_up_to_rounding(orc):
    spec = synthetic.single_block(33, 21)
    b = spec.blocks[0]
    out = orc.tfi(*b.edge_args())
    assert np.abs(out[:, 0] - b.i_min.points).max() < 1e-15
    assert np.abs(out[:, -1] - b.i_max.points).max() < 1e-15
    assert np.abs(out[0, :] - b.j_min.points).max() < 1e-15
    assert np.abs(out[-1, :] - b.j_max.points).max() < 1e-15
    assert not np.isnan(out).any()


def test_tfi_rejects_bad_clustering_and_corners(orc):
    spec = synthetic.single_block(9, 7)
    args = [a.copy() for a in spec.blocks[0].edge_args()]
    bad = [a.copy() for a in args]
    bad[4][0] = 0.1  # s1[0] != 0 (tfi.zig:135)
    with pytest.raises(orc.OracleError):
        orc.tfi(*bad)
    bad = [a.copy() for a in args]
    bad[2][0, 0] += 1e-3  # x_j_min[0] != x_i_min[0] (tfi.zig:150-153)
    with pytest.raises(orc.OracleError):
        orc.tfi(*bad)


# ------------------------------------------------------------------------------------ Krylov solvers
def _umfpack_known_answer():
    """The 5x5 system of the reference's own UMFPACK test (umfpack.zig:71-97), given there in CSC form."""
    Ap = [0, 2, 5, 9, 10, 12]
    Ai = [0, 1, 0, 2, 4, 1, 2, 3, 4, 2, 1, 4]
    Ax = [2.0, 3.0, 3.0, -1.0, 4.0, 4.0, -3.0, 1.0, 2.0, 2.0, 6.0, 1.0]
    b = np.array([8.0, 45.0, -3.0, 3.0, 19.0])
    A = sp.csc_matrix((Ax, Ai, Ap), shape=(5, 5)).tocsr()
    A.sort_indices()
    return A, b


@pytest.mark.parametrize("solver,precond", [("gmres", "ilu0"), ("gmres", "diagonal"), ("bicgstab", "ilu0"), ("bicgstab", "diagonal")])
def test_krylov_known_answer_umfpack_5x5(orc, solver, precond):
    A, b = _umfpack_known_answer()
    x, st = orc.csr_solve(A.indptr, A.indices, A.data, b, opts=orc.options(solver=solver, preconditioner=precond, rtol=1e-14, atol=1e-14))
    assert np.abs(x - np.arange(1.0, 6.0)).max() < 1e-9, (x, st)


def test_krylov_matches_scipy_on_random_sparse(orc):
    rng = np.random.default_rng(7)
    n = 200
    A = sp.random(n, n, density=0.03, random_state=rng, format="csr") + sp.diags(np.full(n, 4.0))
    A = A.tocsr()
    A.sort_indices()
    b = rng.standard_normal(n)
    ref = spla.splu(A.tocsc()).solve(b)
    for solver in ("gmres", "bicgstab"):
        x, st = orc.csr_solve(A.indptr, A.indices, A.data, b, opts=orc.options(solver=solver, preconditioner="ilu0", rtol=1e-13, atol=1e-13))
        assert np.abs(x - ref).max() < 1e-9, (solver, st)


# ------------------------------------------------------------------------- assembled system vs scipy
def _csr(system):
    p, i, v, rx, ry = system.csr()
    return sp.csr_matrix((v, i, p), shape=(system.dof, system.dof)), rx, ry


@pytest.mark.parametrize("name,args", [("cascade", (4, 2, 13, 9)), ("cascade", (2, 2, 12, 9)), ("cascade", (4, 1, 10, 14)), ("single", (21, 17))])
def test_one_picard_step_equals_direct_solve(orc, name, args):
    """The oracle's tight Krylov solve of the assembled system equals an independent sparse-LU solve (what the
    reference's `umfpack` option computes, umfpack.zig:42-53)."""
    spec = synthetic.cascade(*args) if name == "cascade" else synthetic.single_block(*args)
    mesh = synthetic.materialize(spec, orc.tfi)
    ref = mesh.copy()
    S = orc.System(ref, orc.tight_options())
    S.fill(0)
    p, i, v, _, _ = S.csr()
    # reference invariants: ascending columns in every row (smooth.zig:679-687), square system, one row per node
    for r in range(S.dof):
        assert np.all(np.diff(i[p[r]:p[r + 1]]) > 0)
    sol = []
    for y_mode in (False, True):
        S.fill_specific(y_mode)
        A, rx, ry = _csr(S)
        sol.append(spla.splu(A.tocsc()).solve(ry if y_mode else rx))
    direct = np.stack(sol, axis=1)
    orc.smooth_mesh(mesh, 1, orc.tight_options())
    got = np.concatenate([b.points.reshape(-1, 2) for b in mesh.blocks])
    assert np.abs(got - direct).max() < 1e-11


def test_uniform_cartesian_grid_is_a_fixed_point(orc):
    ni, nj = 12, 9
    x, y = np.meshgrid(np.linspace(0, 1.1, ni), np.linspace(0, 0.8, nj), indexing="ij")
    mesh = Mesh([Block2d(np.stack([x, y], axis=-1))], ["b"], [], [])
    ref = mesh.copy()
    st = orc.smooth_mesh(mesh, 2, orc.tight_options())
    assert max_diff(mesh, ref) < 1e-14 and st["last_max_update"] < 1e-14


def test_affine_grid_is_a_fixed_point_and_smoothing_is_translation_equivariant(orc):
    spec = synthetic.cascade(2, 2, 10, 8)
    a = synthetic.materialize(spec, orc.tfi)
    b = a.copy()
    shift = np.array([0.25, -0.125])  # exactly representable, keeps the 1e-15 coincidence check valid
    for blk in b.blocks:
        blk.points += shift
    orc.smooth_mesh(a, 3, orc.tight_options())
    orc.smooth_mesh(b, 3, orc.tight_options())
    for x, y in zip(a.blocks, b.blocks):
        assert np.abs((x.points + shift) - y.points).max() < 1e-11


# -------------------------------------------------------------------------------- topology / kinds
def test_t106_classification_counts(orc):
    spec, z, meta = load_fixture("t106_white")
    mesh = synthetic.materialize(spec, orc.tfi)
    assert [b.size for b in mesh.blocks] == [(221, 41), (121, 41), (11, 41), (11, 51), (121, 41), (161, 11), (21, 91), (11, 131)]
    assert mesh.num_nodes() == 25118
    for k, b in enumerate(mesh.blocks):  # golden TFI
        assert np.array_equal(b.points, z[f"tfi_b{k}"])
    S = orc.System(mesh)
    kinds = np.bincount(S.kinds(), minlength=5)
    assert kinds.sum() == sum(2 * (b.size[0] + b.size[1] - 2) for b in mesh.blocks)
    assert kinds[3] == len(S.junctions()) == 12
    assert kinds[1] == sum(c.len() - 2 for c in mesh.connections)  # one smoothed node per interior interface node
    assert kinds[4] == 220  # inlet 91 + outlet 131 minus the two corners taken by periodic connections
    for j in S.junctions():
        assert len(j["ids"]) <= 4 and len(j["stencil"]) <= 6 and j["ids"] == sorted(j["ids"])


def test_connection_mismatch_is_reported(orc):
    spec = synthetic.cascade(2, 2, 10, 8)
    mesh = synthetic.materialize(spec, orc.tfi)
    mesh.blocks[2].points[0, 3, 0] += 1e-12  # block (1,0), j_min side: > 1e-15 (smooth.zig:221)
    with pytest.raises(orc.OracleError, match="non matching"):
        orc.smooth_mesh(mesh, 1)


def test_single_block_without_connections_smooths(orc):
    """Deviation D1: the reference underflows on a mesh with zero connections (smooth.zig:1364)."""
    mesh = synthetic.materialize(synthetic.single_block(25, 19), orc.tfi)
    ref = mesh.copy()
    st = orc.smooth_mesh(mesh, 3, orc.tight_options())
    assert st["not_converged"] == 0 and 1e-6 < max_diff(mesh, ref) < 0.1
    for a, b in zip(mesh.blocks, ref.blocks):  # all block-boundary nodes are fixed
        for sl in (np.s_[0, :], np.s_[-1, :], np.s_[:, 0], np.s_[:, -1]):
            assert np.array_equal(a.points[sl], b.points[sl])


def test_golden_t106_white_is_reproduced(orc):
    """The committed golden mesh is what the oracle computes from the committed inputs (10 outer iterations, White)."""
    spec, z, meta = load_fixture("t106_laplace")
    mesh = synthetic.materialize(spec, orc.tfi)
    orc.smooth_mesh(mesh, meta["iterations"], orc.tight_options(max_iters=100000))
    err = max(float(np.abs(b.points - z[f"smooth_b{k}"]).max()) for k, b in enumerate(mesh.blocks))
    assert err < 1e-12


def test_block_to_soa_is_the_cgns_layout(orc):
    """cgns.zig:69-101: buffer[j*ni + i] = block(i, j) -- i fastest."""
    ni, nj = 7, 5
    a = np.arange(ni * nj * 2, dtype=np.float64).reshape(ni, nj, 2)
    x, y = orc.block_to_soa(a)
    assert np.array_equal(x, a[:, :, 0].T.ravel()) and np.array_equal(y, a[:, :, 1].T.ravel())


def test_edge_discretisation_oracle_matches_the_host_mirror(orc):
    """clustering.create + Curve.interpolate (discrete.zig:17-31): the C restatement and the host-side mirror agree bit for
    bit (same libm), incl. the reference's straight-line known answer for the spline (spline.zig:235-304)."""
    from turbomesh_b200.clustering import Roberts, SingleHyperbolicClustering, Uniform
    from inputgen.geometry import Line
    from inputgen.spline import FittingSpline

    for n in (2, 5, 41, 200):
        assert np.array_equal(orc.clustering("uniform", n), Uniform().compute(n))
        assert np.array_equal(orc.clustering("roberts", n, alpha=0.5, beta=1.03), Roberts(0.5, 1.03).compute(n))
    for n, ds in ((5, 0.2), (41, 0.01), (131, 1e-3)):
        c = orc.clustering("single_hyperbolic_clustering", n, delta_s=ds)
        assert np.array_equal(c, SingleHyperbolicClustering(ds).compute(n)) and c[0] == 0.0 and c[-1] == 1.0
    t = np.linspace(0.0, 1.0, 37)
    sp = FittingSpline(np.stack([t, 0.1 * np.sin(3 * t) + 0.2 * t ** 2], axis=1))
    u = Roberts(0.5, 1.1).compute(60)
    got = orc.spline_interpolate(sp.params, sp.points, sp.second_derivs[0], sp.second_derivs[1], sp.sample_arc, sp.total_length, u)
    assert np.array_equal(got, sp.interpolate(u))
    line = FittingSpline(np.array([[0.0, 0.0], [1.0, 2.0], [2.0, 4.0], [4.0, 8.0]]))   # spline.zig:235-260: a straight line is reproduced
    uu = Uniform().compute(9)
    pts = orc.spline_interpolate(line.params, line.points, line.second_derivs[0], line.second_derivs[1], line.sample_arc, line.total_length, uu)
    assert np.allclose(pts, np.stack([4.0 * uu, 8.0 * uu], axis=1), atol=1e-9)
    ln = Line((0.1, 0.2), (1.3, -0.4))
    assert np.array_equal(orc.line_interpolate(ln.start, ln.end, uu), ln.interpolate(uu))


def test_viewer_buffers_oracle(orc):
    """gui/lib.zig:227-318: f32 points in block order with ranges, line indices per block (along j, then along i)."""
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=(4, 3, 2)), rng.normal(size=(3, 5, 2)) + 2.0
    pts, rx, ry, idx = orc.viewer_buffers([a, b])
    ref = np.concatenate([a.reshape(-1, 2), b.reshape(-1, 2)]).astype(np.float32)
    assert np.array_equal(pts, ref.ravel())
    assert rx[0] == ref[:, 0].min() and rx[1] == ref[:, 0].max() and ry[0] == ref[:, 1].min() and ry[1] == ref[:, 1].max()
    want, off = [], 0
    for ni, nj in ((4, 3), (3, 5)):
        want += [v for i in range(ni) for j in range(nj - 1) for v in (off + i * nj + j, off + i * nj + j + 1)]
        want += [v for j in range(nj) for i in range(ni - 1) for v in (off + i * nj + j, off + (i + 1) * nj + j)]
        off += ni * nj
    assert np.array_equal(idx, np.array(want, dtype=np.uint32))
    neg = orc.viewer_buffers([-np.abs(a) - 1.0])                      # all coordinates negative: the maxima keep their start value
    assert neg[1][1] == np.float32(1.1754944e-38) and neg[2][1] == np.float32(1.1754944e-38)


# -------------------------------------------------------------------------------- more identities the operator must obey
def _smoothed_single_block(orc, points, iterations=3):
    mesh = Mesh([Block2d(np.array(points, dtype=np.float64, order="C", copy=True))], ["b"], [], [])  # smoothing is in place
    orc.smooth_mesh(mesh, iterations, orc.tight_options())
    return mesh.blocks[0].points


def test_smoothing_is_scale_rotation_and_reflection_equivariant(orc):
    """The Winslow operator only sees g11, g22, g12 (smooth.zig:192-215): scaling, rotating or mirroring the mesh commutes
    with smoothing when no node slides along an axis.  The maps below are exact in binary floating point."""
    base = orc.tfi(*synthetic.single_block(21, 17).blocks[0].edge_args())
    ref = _smoothed_single_block(orc, base)
    assert np.abs(ref - base).max() > 1e-4  # the test smooths something
    maps = {
        "scale x4": lambda p: 4.0 * p,
        "rotate 90": lambda p: np.stack([-p[..., 1], p[..., 0]], axis=-1),
        "mirror x": lambda p: np.stack([-p[..., 0], p[..., 1]], axis=-1),
        "swap i and j": None,
    }
    for name, f in maps.items():
        if f is None:  # the same geometry traversed with the index directions exchanged
            got = _smoothed_single_block(orc, np.ascontiguousarray(base.transpose(1, 0, 2))).transpose(1, 0, 2)
            want, scale = ref, 1.0
        else:
            got, want, scale = _smoothed_single_block(orc, f(base)), f(ref), 4.0 if name.startswith("scale") else 1.0
        assert np.abs(got - want).max() <= 1e-11 * scale, name


@pytest.mark.parametrize("cut", ["i", "j"])
def test_a_block_split_by_a_connection_smooths_like_the_unsplit_block(orc, cut):
    """Interface rows (smooth.zig:994-1105) are interior rows written across two blocks: cutting a block in two along a
    grid line and joining the halves with an ordinary connection must not change the smoothed mesh."""
    from turbomesh_b200.boundary import Connection, Range, Side

    base = orc.tfi(*synthetic.single_block(25, 19).blocks[0].edge_args())
    ref = _smoothed_single_block(orc, base, iterations=4)
    if cut == "i":   # halves share the grid line i = 12: side j_max of the first, j_min of the second (boundary.zig:34-51)
        a, b = base[:13].copy(), base[12:].copy()
        conn = Connection((Range(0, Side.j_max, 0, base.shape[1] - 1), Range(1, Side.j_min, 0, base.shape[1] - 1)), None)
    else:            # halves share the grid line j = 9: side i_max of the first, i_min of the second
        a, b = np.ascontiguousarray(base[:, :10]), np.ascontiguousarray(base[:, 9:])
        conn = Connection((Range(0, Side.i_max, 0, base.shape[0] - 1), Range(1, Side.i_min, 0, base.shape[0] - 1)), None)
    mesh = Mesh([Block2d(a), Block2d(b)], ["a", "b"], [conn], [])
    orc.smooth_mesh(mesh, 4, orc.tight_options())
    if cut == "i":
        got = np.concatenate([mesh.blocks[0].points, mesh.blocks[1].points[1:]], axis=0)
        seam = np.abs(mesh.blocks[0].points[-1] - mesh.blocks[1].points[0]).max()
    else:
        got = np.concatenate([mesh.blocks[0].points, mesh.blocks[1].points[:, 1:]], axis=1)
        seam = np.abs(mesh.blocks[0].points[:, -1] - mesh.blocks[1].points[:, 0]).max()
    assert seam <= 1e-14
    assert np.abs(got - ref).max() <= 1e-11


def test_moving_the_seam_of_a_periodic_block_does_not_change_the_mesh(orc):
    """A channel that is periodic across j (same-block connection i_min -> i_max with a periodicity vector, the only
    same-block form the reference supports, smooth.zig:522-559): where the seam lies is arbitrary, so rolling the columns
    by k and smoothing must give the rolled mesh -- the periodic interface rows with their shifted neighbours and
    right-hand sides (smooth.zig:1029-1061) are just interior rows."""
    from turbomesh_b200.boundary import Connection, Range, Side

    ni, nj, height = 17, 14, 0.5  # nj - 1 distinct columns
    s = np.linspace(0.0, 1.0, ni)[:, None]
    col = np.arange(nj - 1)[None, :]

    def channel(columns):
        t = columns / (nj - 1.0)
        x = s + 0.03 * np.sin(2 * np.pi * t) * np.sin(np.pi * s) + 0.02 * np.sin(np.pi * s) * np.cos(4 * np.pi * t)
        y = height * t + 0.04 * np.sin(2 * np.pi * s) + 0.015 * np.sin(np.pi * s) * np.sin(2 * np.pi * t)
        return np.stack([x + 0 * t, y], axis=-1)

    def smoothed(k):
        """Columns k, k+1, ..., k+nj-2 and, as last column, the periodic image of the first."""
        pts = channel(col + k)
        pts = np.concatenate([pts, pts[:, :1] + np.array([0.0, height])], axis=1)
        conn = Connection((Range(0, Side.i_min, 0, ni - 1), Range(0, Side.i_max, 0, ni - 1)), (0.0, height))
        mesh = Mesh([Block2d(pts.copy())], ["ring"], [conn], [])  # smoothing is in place
        orc.smooth_mesh(mesh, 4, orc.tight_options())
        out = mesh.blocks[0].points
        assert np.abs(out[:, -1] - out[:, 0] - np.array([0.0, height])).max() <= 1e-14
        return out[:, :-1], pts[:, :-1]

    ref, start = smoothed(0)
    assert np.abs(ref - start).max() > 1e-4
    for k in (1, 5, nj - 2):
        got, _ = smoothed(k)
        # column c of the rolled mesh is column (c + k) of the reference, lifted by one period where it wrapped around
        want = np.roll(ref, -k, axis=1).copy()
        want[:, (nj - 1) - k:, 1] += height
        assert np.abs(got - want).max() <= 1e-11, k
