"""First GPU parity tests: CUDA path through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from turbomesh_b200 import synthetic

pytestmark = pytest.mark.gpu


def _max_diff(a, b):
    return max(float(np.abs(x.points - y.points).max()) for x, y in zip(a.blocks, b.blocks))


@pytest.mark.parametrize("shape", [(3, 3), (5, 7), (129, 33), (64, 257), (300, 131)])
def test_tfi_bit_exact_single_block(orc, gpu_lib, shape):
    from turbomesh_b200 import smoothing

    spec = synthetic.single_block(*shape)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    assert np.array_equal(gpu.blocks[0].points, cpu.blocks[0].points)


def test_tfi_bit_exact_different_clusterings(orc, gpu_lib):
    from turbomesh_b200 import smoothing
    from turbomesh_b200.clustering import Roberts, SingleHyperbolicClustering, Uniform

    ni, nj = 77, 53
    s1, s2 = Roberts(0.5, 1.03).compute(ni), Uniform().compute(ni)
    t1, t2 = SingleHyperbolicClustering(0.01).compute(nj), Roberts(0.5, 1.2).compute(nj)
    for c in (s1, s2, t1, t2):
        c[0], c[-1] = 0.0, 1.0
    xi_min = np.stack([s1, 0.1 * np.sin(3 * s1)], axis=1)
    xi_max = np.stack([s2 * 1.1 - 0.05, 1 + 0.1 * np.cos(2 * s2)], axis=1)
    xj_min = np.stack([xi_min[0, 0] + (xi_max[0, 0] - xi_min[0, 0]) * t1, xi_min[0, 1] + (xi_max[0, 1] - xi_min[0, 1]) * t1 + 0.05 * np.sin(np.pi * t1)], axis=1)
    xj_max = np.stack([xi_min[-1, 0] + (xi_max[-1, 0] - xi_min[-1, 0]) * t2, xi_min[-1, 1] + (xi_max[-1, 1] - xi_min[-1, 1]) * t2], axis=1)
    xj_min[0], xj_min[-1], xj_max[0], xj_max[-1] = xi_min[0], xi_max[0], xi_min[-1], xi_max[-1]
    a = smoothing.tfi_block(xi_min, xi_max, xj_min, xj_max, s1, s2, t1, t2)
    b = orc.tfi(xi_min, xi_max, xj_min, xj_max, s1, s2, t1, t2)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("args", [(2, 2, 12, 9), (4, 2, 33, 17), (4, 1, 20, 30), (8, 8, 16, 12)])
def test_picard_bicgstab_matches_oracle_cascade(orc, gpu_lib, args):
    from turbomesh_b200 import smoothing

    spec = synthetic.cascade(*args)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    assert _max_diff(gpu, cpu) == 0.0
    st = smoothing.smooth_mesh(gpu, 4, smoothing.CudaSolver.tight())
    orc.smooth_mesh(cpu, 4, orc.tight_options())
    chord = 1.0
    assert _max_diff(gpu, cpu) <= 1e-9 * chord, st


def test_picard_bicgstab_matches_oracle_single_block(orc, gpu_lib):
    from turbomesh_b200 import smoothing

    spec = synthetic.single_block(65, 49)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    smoothing.smooth_mesh(gpu, 5, smoothing.CudaSolver.tight())
    orc.smooth_mesh(cpu, 5, orc.tight_options())
    assert _max_diff(gpu, cpu) <= 1e-9


def test_relax_converges_to_picard_fixed_point(orc, gpu_lib):
    """Nonlinear Jacobi sweeps and the Picard iteration share their fixed point."""
    from turbomesh_b200 import smoothing

    spec = synthetic.cascade(2, 2, 12, 9)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    st = smoothing.smooth_mesh(gpu, 400, smoothing.CudaSolver(method="relax", sweeps_per_iteration=50, omega=0.9, stop_max_update=1e-14))
    orc.smooth_mesh(cpu, 30, orc.tight_options())
    assert _max_diff(gpu, cpu) <= 1e-9, st


@pytest.mark.parametrize("shape", [(129, 129), (130, 75), (257, 64)])
def test_multigrid_converges_to_the_oracle_fixed_point(orc, gpu_lib, shape):
    """FAS multigrid (time-to-converged path) reaches the fixed point of the reference's Picard iteration; the levels are
    non-nested for (130, 75), semi-coarsened for (257, 64)."""
    from turbomesh_b200 import smoothing

    spec = synthetic.single_block(*shape)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    st = smoothing.smooth_mesh(gpu, 80, smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-13))
    orc.smooth_mesh(cpu, 40, orc.tight_options())
    assert st["last_max_update"] <= 1e-13 and st["outer_iterations"] < 80, st
    assert _max_diff(gpu, cpu) <= 1e-9, st


@pytest.mark.parametrize("args", [(2, 2, 33, 17), (4, 2, 33, 17), (4, 1, 17, 33), (4, 4, 17, 9)])
def test_multiblock_multigrid_converges_to_the_oracle_fixed_point(orc, gpu_lib, args):
    """Multi-block FAS multigrid (interfaces, periodic rows, junctions, plate ends, sliding inlet / outlet on every level)
    reaches the fixed point of the reference's Picard iteration."""
    from turbomesh_b200 import smoothing

    spec = synthetic.cascade(*args)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    st = smoothing.smooth_mesh(gpu, 80, smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-13))
    orc.smooth_mesh(cpu, 40, orc.tight_options())
    assert st["last_max_update"] <= 1e-13 and st["outer_iterations"] < 80, st
    assert _max_diff(gpu, cpu) <= 1e-9, st


def test_multiblock_multigrid_without_coarsenable_direction(orc, gpu_lib):
    """11 x 8 intervals per block: only j can be halved (three times); the cycle degrades gracefully and keeps the fixed point."""
    from turbomesh_b200 import smoothing

    spec = synthetic.cascade(2, 2, 12, 9)
    gpu = synthetic.materialize(spec, smoothing.tfi_block)
    cpu = synthetic.materialize(spec, orc.tfi)
    st = smoothing.smooth_mesh(gpu, 400, smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-13))
    orc.smooth_mesh(cpu, 40, orc.tight_options())
    assert st["last_max_update"] <= 1e-13, st
    assert _max_diff(gpu, cpu) <= 1e-9, st


def test_multigrid_rejects_the_white_control_function(orc, gpu_lib):
    from turbomesh_b200 import _lib, smoothing

    mesh = synthetic.materialize(synthetic.cascade(2, 2, 17, 9), orc.tfi)
    with pytest.raises(_lib.TurbomeshGpuError) as e:
        smoothing.smooth_mesh(mesh, 2, smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=2), smoothing.White(1e-4))
    assert e.value.code == _lib.TM_ERR_UNSUPPORTED


@pytest.mark.parametrize("shape,iterations,sweeps", [((700, 160), 2, 12), ((300, 67), 1, 30), ((1500, 40), 3, 5)])
def test_streamed_host_smoothing_is_bit_identical(gpu_lib, monkeypatch, shape, iterations, sweeps):
    """tm_smooth_mesh streams a large single block through the device in row chunks, copies overlapped with the sweeps
    (streamed.inl); every node gets the arithmetic of the resident path, so the meshes are equal bit for bit."""
    from turbomesh_b200 import smoothing
    from turbomesh_b200.discrete import Block2d, Mesh

    base = synthetic.materialize(synthetic.single_block(*shape), smoothing.tfi_block).blocks[0].points
    solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=0.9)

    def run():
        mesh = Mesh()
        mesh.add_block("block", Block2d(base.copy()))
        return mesh.blocks[0].points, smoothing.smooth_mesh(mesh, iterations, solver)

    monkeypatch.setenv("TM_STREAM", "0")
    resident, st_r = run()
    monkeypatch.setenv("TM_STREAM", "1")
    monkeypatch.setenv("TM_STREAM_MIN_NODES", "0")
    plan = smoothing.stream_plan(shape[0], shape[1], iterations * sweeps)
    assert plan is not None
    streamed, st_s = run()
    assert st_r["streamed_chunks"] == 0 and st_s["streamed_chunks"] == len(plan[1]) >= 2
    assert np.array_equal(streamed, resident)
    assert not np.array_equal(streamed, base)
    assert st_s["last_max_update"] == st_r["last_max_update"]
    assert st_s["last_sumsq_x"] == pytest.approx(st_r["last_sumsq_x"], rel=1e-12) and st_s["last_sumsq_y"] == pytest.approx(st_r["last_sumsq_y"], rel=1e-12)
    for key in ("outer_iterations", "inner_iterations", "operator_applications", "nodes", "converged"):
        assert st_s[key] == st_r[key], key


@pytest.mark.parametrize("cut", ["i", "j"])
def test_a_block_split_by_a_connection_smooths_like_the_unsplit_block(orc, gpu_lib, cut):
    """Interface rows are interior rows written across two blocks (smooth.zig:994-1105): cutting a block in two along a
    grid line and joining the halves with an ordinary connection must not change the smoothed mesh -- an identity that
    needs no oracle (tests/test_oracle_cpu.py holds the same test for the oracle)."""
    from turbomesh_b200 import smoothing
    from turbomesh_b200.boundary import Connection, Range, Side
    from turbomesh_b200.discrete import Block2d, Mesh

    base = smoothing.tfi_block(*synthetic.single_block(49, 37).blocks[0].edge_args())
    whole = Mesh([Block2d(base.copy())], ["b"], [], [])
    smoothing.smooth_mesh(whole, 4, smoothing.CudaSolver.tight())
    ref = whole.blocks[0].points
    assert np.abs(ref - base).max() > 1e-5
    if cut == "i":
        a, b = base[:25].copy(), base[24:].copy()
        conn = Connection((Range(0, Side.j_max, 0, base.shape[1] - 1), Range(1, Side.j_min, 0, base.shape[1] - 1)), None)
    else:
        a, b = np.ascontiguousarray(base[:, :19]), np.ascontiguousarray(base[:, 18:])
        conn = Connection((Range(0, Side.i_max, 0, base.shape[0] - 1), Range(1, Side.i_min, 0, base.shape[0] - 1)), None)
    halves = Mesh([Block2d(a), Block2d(b)], ["a", "b"], [conn], [])
    smoothing.smooth_mesh(halves, 4, smoothing.CudaSolver.tight())
    pa, pb = halves.blocks[0].points, halves.blocks[1].points
    if cut == "i":
        got, seam = np.concatenate([pa, pb[1:]], axis=0), np.abs(pa[-1] - pb[0]).max()
    else:
        got, seam = np.concatenate([pa, pb[:, 1:]], axis=1), np.abs(pa[:, -1] - pb[:, 0]).max()
    assert seam == 0.0                       # connected copies are exact copies
    assert np.abs(got - ref).max() <= 1e-9
    # the relaxation path: same identity sweep by sweep (to rounding: the halves sum their partial results differently)
    relax = smoothing.CudaSolver(method="relax", sweeps_per_iteration=25, omega=0.9)
    whole = Mesh([Block2d(base.copy())], ["b"], [], [])
    smoothing.smooth_mesh(whole, 1, relax)
    fresh = (base[:25].copy(), base[24:].copy()) if cut == "i" else (np.ascontiguousarray(base[:, :19]), np.ascontiguousarray(base[:, 18:]))
    halves = Mesh([Block2d(fresh[0]), Block2d(fresh[1])], ["a", "b"], [conn], [])
    smoothing.smooth_mesh(halves, 1, relax)
    pa, pb = halves.blocks[0].points, halves.blocks[1].points
    got = np.concatenate([pa, pb[1:]], axis=0) if cut == "i" else np.concatenate([pa, pb[:, 1:]], axis=1)
    assert np.abs(got - whole.blocks[0].points).max() <= 1e-13


def test_moving_the_seam_of_a_periodic_block_does_not_change_the_mesh(gpu_lib):
    """A channel periodic across j (same-block connection i_min -> i_max with a periodicity vector): rolling the columns by
    k and smoothing gives the rolled mesh -- the periodic interface rows with their shifted neighbours are interior rows
    (tests/test_oracle_cpu.py holds the same identity for the oracle)."""
    from turbomesh_b200 import smoothing
    from turbomesh_b200.boundary import Connection, Range, Side
    from turbomesh_b200.discrete import Block2d, Mesh

    ni, nj, height = 33, 26, 0.5
    s = np.linspace(0.0, 1.0, ni)[:, None]
    col = np.arange(nj - 1)[None, :]

    def smoothed(k, solver, iterations):
        t = (col + k) / (nj - 1.0)
        x = s + 0.03 * np.sin(2 * np.pi * t) * np.sin(np.pi * s) + 0.02 * np.sin(np.pi * s) * np.cos(4 * np.pi * t)
        y = height * t + 0.04 * np.sin(2 * np.pi * s) + 0.015 * np.sin(np.pi * s) * np.sin(2 * np.pi * t)
        pts = np.stack([x, y], axis=-1)
        pts = np.concatenate([pts, pts[:, :1] + np.array([0.0, height])], axis=1)
        conn = Connection((Range(0, Side.i_min, 0, ni - 1), Range(0, Side.i_max, 0, ni - 1)), (0.0, height))
        mesh = Mesh([Block2d(pts.copy())], ["ring"], [conn], [])
        smoothing.smooth_mesh(mesh, iterations, solver)
        out = mesh.blocks[0].points
        assert np.abs(out[1:-1, -1] - (out[1:-1, 0] + np.array([0.0, height]))).max() == 0.0   # copies are exact copies
        return out[:, :-1], pts[:, :-1]

    for solver, iterations, tol in ((smoothing.CudaSolver.tight(), 4, 1e-9), (smoothing.CudaSolver(method="relax", sweeps_per_iteration=30, omega=0.9), 1, 1e-13)):
        ref, start = smoothed(0, solver, iterations)
        assert np.abs(ref - start).max() > 1e-4
        for k in (1, 9, nj - 2):
            got, _ = smoothed(k, solver, iterations)
            want = np.roll(ref, -k, axis=1).copy()
            want[:, (nj - 1) - k:, 1] += height
            assert np.abs(got - want).max() <= tol, (solver.method, k)
