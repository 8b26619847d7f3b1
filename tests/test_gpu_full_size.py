"""Parity at BASELINE.json's full single-GPU size (config 3: one 8192 x 8192 fp64 block, 67 M nodes).

The oracle's Krylov path cannot run at this size (~30 GB of work vectors), so the smoother is
checked through properties that do not depend on the size:

* TFI is a pure function of the edges -> bit-exact against the oracle on the whole block (the oracle's TFI is O(N));
* a relaxation sweep is local (a node's new value depends on its 3 x 3 neighbourhood only) -> windows cut out of the
  full-size field are handed to the ORACLE's own system assembly (``RowCompressedMatrixSystem2d``, smooth.zig:923-992
  through ``oracle.System``), and the damped-Jacobi update computed from its CSR rows must agree with what the kernel
  wrote for the same nodes;
* the converged mesh is a fixed point of the reference's Picard step -> the oracle's rows, assembled on windows of the
  converged mesh, have a residual (in units of length: the Jacobi update) below the stopping criterion;
* a uniform Cartesian grid is a fixed point; the smoother is translation equivariant.
"""
import numpy as np
import pytest

from turbomesh_b200 import synthetic
from turbomesh_b200.discrete import Block2d, Mesh

pytestmark = pytest.mark.gpu

N = 8192
W = 96  # window extent (nodes); windows sit in the corners, on the edges and in the middle of the block
OMEGA = 0.9


def _windows(n, w):
    starts = [0, n // 2 - w // 2, n - w]
    return [(i0, j0) for i0 in starts for j0 in starts] + [(1234, 5), (7000, n - w - 3)]


def _oracle_jacobi_update(orc, window, omega):
    """Damped-Jacobi update of the interior nodes of ``window`` from the oracle's assembled rows: x + w (b - A x)_i / a_ii."""
    mesh = Mesh()
    mesh.add_block("window", Block2d(np.ascontiguousarray(window)))
    sys_ = orc.System(mesh, orc.options(control_function="laplace"))
    sys_.fill(0)
    p, idx, v, rx, ry = sys_.csr()
    sys_.close()
    ni, nj = window.shape[:2]
    flat = window.reshape(-1, 2)
    out = flat.copy()
    rows = (np.arange(1, ni - 1)[:, None] * nj + np.arange(1, nj - 1)[None, :]).ravel()
    for r in rows:
        a, cols = v[p[r]:p[r + 1]], idx[p[r]:p[r + 1]]
        assert len(cols) == 9
        diag = a[cols == r][0]
        res_x = rx[r] - np.dot(a, flat[cols, 0])
        res_y = ry[r] - np.dot(a, flat[cols, 1])
        out[r, 0] = flat[r, 0] + omega * res_x / diag
        out[r, 1] = flat[r, 1] + omega * res_y / diag
    return out.reshape(ni, nj, 2)


@pytest.fixture(scope="module")
def full_block(orc, gpu_lib):
    """(spec, TFI of the 8192^2 block through tm_tfi_block with host buffers)."""
    from turbomesh_b200 import smoothing

    spec = synthetic.single_block(N, N)
    blk = spec.blocks[0]
    return spec, smoothing.tfi_block(*blk.edge_args())


def test_tfi_bit_exact_at_8192x8192(orc, gpu_lib, full_block):
    spec, gpu = full_block
    cpu = orc.tfi(*spec.blocks[0].edge_args())
    assert gpu.shape == (N, N, 2)
    assert np.array_equal(gpu, cpu)


def test_one_sweep_at_8192x8192_matches_the_oracle_rows_on_windows(orc, gpu_lib, full_block):
    from turbomesh_b200 import smoothing

    spec, before = full_block
    mesh = Mesh()
    mesh.add_block("block", Block2d(before.copy()))
    with smoothing.DeviceMesh(mesh) as dm:
        solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=OMEGA)
        dm.begin_smoothing(solver)
        st = dm.smooth(1, solver)
        after = dm.download_block(0)
    assert st["outer_iterations"] == 1
    # fixed boundary nodes never move (smooth.zig:790-796)
    for sl in (np.s_[0, :], np.s_[-1, :], np.s_[:, 0], np.s_[:, -1]):
        assert np.array_equal(after[sl], before[sl])
    worst = 0.0
    for i0, j0 in _windows(N, W):
        want = _oracle_jacobi_update(orc, before[i0:i0 + W, j0:j0 + W], OMEGA)
        got = after[i0:i0 + W, j0:j0 + W]
        worst = max(worst, float(np.abs(got[1:-1, 1:-1] - want[1:-1, 1:-1]).max()))
    # the update itself is ~1e-5; the two evaluations of the row (difference form on the device, a_k x_k on the CPU)
    # differ by rounding of O(eps |x|)
    assert worst <= 5e-15, worst
    assert float(np.abs(after - before).max()) == pytest.approx(st["last_max_update"], rel=1e-12)


def test_converged_8192x8192_mesh_is_a_fixed_point_of_the_oracle_rows(orc, gpu_lib, full_block):
    """time-to-converged path (FAS multigrid) at full size: the stopping criterion of bench.py holds, and the oracle's own
    rows see a converged mesh."""
    from turbomesh_b200 import smoothing

    spec, before = full_block
    mesh = Mesh()
    mesh.add_block("block", Block2d(before.copy()))
    with smoothing.DeviceMesh(mesh) as dm:
        solver = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-10)
        dm.begin_smoothing(solver)
        st = dm.smooth(60, solver)
        after = dm.download_block(0)
    assert st["last_max_update"] <= 1e-10 and st["outer_iterations"] < 60, st
    worst = 0.0
    for i0, j0 in _windows(N, W):
        win = after[i0:i0 + W, j0:j0 + W]
        worst = max(worst, float(np.abs(_oracle_jacobi_update(orc, win, 1.0) - win).max()))
    assert worst <= 1e-10, worst
    assert np.isfinite(after).all()


def test_uniform_grid_is_a_fixed_point_and_sweeps_are_translation_equivariant_at_full_size(gpu_lib, full_block):
    from turbomesh_b200 import smoothing

    solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=5, omega=OMEGA)
    g = np.empty((N, N, 2))
    g[..., 0] = (np.arange(N) / 1024.0)[:, None]  # exactly representable spacings: every difference is exact
    g[..., 1] = (np.arange(N) / 2048.0)[None, :]
    mesh = Mesh()
    mesh.add_block("cartesian", Block2d(g))
    st = smoothing.smooth_mesh(mesh, 1, solver)
    assert st["last_max_update"] == 0.0, st
    del g, mesh

    _, before = full_block
    shift = np.array([0.5, -0.25])  # powers of two: the shifted coordinates stay (almost everywhere) exact
    a, b = Mesh(), Mesh()
    a.add_block("a", Block2d(before.copy()))
    b.add_block("b", Block2d(before + shift))
    smoothing.smooth_mesh(a, 1, solver)
    smoothing.smooth_mesh(b, 1, solver)
    assert float(np.abs(b.blocks[0].points - shift - a.blocks[0].points).max()) <= 1e-14


def test_streamed_and_resident_host_smoothing_agree_bit_for_bit_at_full_size(gpu_lib, full_block, monkeypatch):
    """The bench's end-to-end step (tm_smooth_mesh on the 8192^2 block, 100 sweeps): 8 row chunks streamed through the
    device with the copies overlapped give the mesh of the resident path bit for bit."""
    from turbomesh_b200 import smoothing

    _, before = full_block
    solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=100, omega=OMEGA)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("TM_STREAM", mode)
        mesh = Mesh()
        mesh.add_block("block", Block2d(before.copy()))
        st = smoothing.smooth_mesh(mesh, 1, solver)
        out[mode] = (mesh.blocks[0].points, st)
    assert out["0"][1]["streamed_chunks"] == 0 and out["1"][1]["streamed_chunks"] == 8
    assert np.array_equal(out["0"][0], out["1"][0])
    assert out["0"][1]["last_max_update"] == out["1"][1]["last_max_update"] > 0.0


def _interface_row_updates(orc, a, b, side_a, side_b, periodicity=None):
    """Jacobi updates (b - A x)_r / a_rr of the oracle's `smoothed` rows (smooth.zig:994-1105) of a two-block window mesh
    whose blocks `a` (side_a) and `b` (side_b) are joined over their whole common side."""
    from turbomesh_b200.boundary import Connection, Range

    mesh = Mesh()
    mesh.add_block("a", Block2d(np.ascontiguousarray(a)))
    mesh.add_block("b", Block2d(np.ascontiguousarray(b)))
    n, nj = a.shape[0], a.shape[1]
    mesh.connections.append(Connection((Range(0, side_a, 0, n - 1), Range(1, side_b, 0, n - 1)), periodicity))
    sys_ = orc.System(mesh, orc.options(control_function="laplace"))
    sys_.fill(0)
    p, idx, v, rx, ry = sys_.csr()
    sys_.close()
    flat = np.concatenate([a.reshape(-1, 2), b.reshape(-1, 2)])
    j = 0 if int(side_a) == 0 else nj - 1
    worst = 0.0
    for i in range(1, n - 1):
        r = i * nj + j
        cols, val = idx[p[r]:p[r + 1]], v[p[r]:p[r + 1]]
        assert len(cols) == 9
        diag = val[cols == r][0]
        worst = max(worst, abs((rx[r] - val @ flat[cols, 0]) / diag), abs((ry[r] - val @ flat[cols, 1]) / diag))
    return worst


def test_config4_block_column_at_full_size_against_the_oracle_rows(orc, gpu_lib):
    """One GPU's share of config 4 (8 blocks of 4097 x 2049 = 67 M nodes: interfaces, a periodic pair, the plate's walls,
    sliding inlet / outlet): TFI bit-exact, and on the multigrid-converged mesh the oracle's own interior AND interface
    rows, assembled on windows cut out of the full-size blocks, are satisfied; copies, walls and sliding nodes hold their
    defining relations."""
    from turbomesh_b200 import smoothing
    from turbomesh_b200.boundary import Side

    ni, nj, height = 4097, 2049, 0.5
    spec = synthetic.cascade(1, 8, ni, nj, length=1.0 / 8.0, ay=0.015 / 8.0)
    solver = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-10)
    with smoothing.DeviceMesh(spec, upload=False) as dm:
        for k, b in enumerate(spec.blocks):
            dm.tfi_block(k, *b.edge_args())
        tfi = {k: dm.download_block(k) for k in (0, 3)}
        dm.begin_smoothing(solver)
        st = dm.smooth(100, solver)
        got = {k: dm.download_block(k) for k in (0, 1, 3, 4, 7)}
    for k in tfi:
        assert np.array_equal(tfi[k], orc.tfi(*spec.blocks[k].edge_args()))
    assert st["last_max_update"] <= 1e-10 and st["outer_iterations"] < 100, st
    tol = 2e-10
    wi, wj = 64, 24
    for i0 in (1, ni // 2, ni - wi - 1):
        rows = slice(i0, i0 + wi)
        # blocks 0 | 1: ordinary interface (side i_max of block 0 = its last j line, side i_min of block 1 = its first)
        assert _interface_row_updates(orc, got[0][rows, -wj:], got[1][rows, :wj], Side.i_max, Side.i_min) <= tol
        # blocks 0 | 7: the periodic pair, x0 + p == x7
        assert _interface_row_updates(orc, got[0][rows, :wj], got[7][rows, -wj:], Side.i_min, Side.i_max, (0.0, height)) <= tol
        # interior rows
        win = got[1][rows, 1000:1000 + wi]
        assert float(np.abs(_oracle_jacobi_update(orc, win, 1.0) - win).max()) <= tol
    inner = slice(1, ni - 1)
    assert np.array_equal(got[0][inner, -1], got[1][inner, 0])                                   # connected copies are exact copies
    assert float(np.abs(got[0][inner, 0] + np.array([0.0, height]) - got[7][inner, -1]).max()) <= 1e-15
    assert np.array_equal(got[3][inner, -1], tfi[3][inner, -1])                                  # the plate's lower face: fixed wall
    assert np.array_equal(got[4][inner, 0], orc.tfi(*spec.blocks[4].edge_args())[inner, 0])      # ... and its upper face
    mid = slice(1, nj - 1)
    assert np.array_equal(got[1][0, mid, 0], np.zeros(nj - 2))                                   # inlet: x stays, y follows its inner neighbour
    assert float(np.abs(got[1][0, mid, 1] - got[1][1, mid, 1]).max()) <= tol
