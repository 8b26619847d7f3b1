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
