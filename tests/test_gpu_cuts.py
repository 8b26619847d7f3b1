"""Config 5: a batch of independent 2D cuts (scaled T106 block sets) smoothed as one mesh, against per-cut oracle runs."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import load_fixture  # noqa: E402

from turbomesh_b200 import synthetic  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("control", ["laplace", "white"])
def test_batch_of_cuts_matches_per_cut_oracle(gpu_lib, orc, control):
    from turbomesh_b200 import smoothing

    base, z, meta = load_fixture("t106_white")
    scales = [1.0, 1.0625, 1.125]
    iterations = 3
    batch, groups = synthetic.batch_of_cuts(base, scales)
    nb = len(base.blocks)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"]) if control == "white" else smoothing.Laplace()
    with smoothing.DeviceMesh(batch, upload=False) as dm:
        for k, b in enumerate(batch.blocks):
            dm.tfi_block(k, *b.edge_args())
        if control == "white":
            dm.set_white_groups(groups)
        sol = smoothing.CudaSolver.tight()
        dm.begin_smoothing(sol, cf)
        st = dm.smooth(iterations, sol, cf)
        got = [dm.download_block(k) for k in range(len(batch.blocks))]
    assert st["converged"] == 1 and st["nodes"] == len(scales) * 25118
    for c, sc in enumerate(scales):
        cut = synthetic.Mesh(names=list(base.names))
        single, _ = synthetic.batch_of_cuts(base, [sc])
        cpu = synthetic.materialize(single, orc.tfi)
        kw = dict(control_function=control)
        if control == "white":
            kw.update(ds_target=meta["ds_target"], theta_target=meta["theta_target"])
        orc.smooth_mesh(cpu, iterations, orc.tight_options(max_iters=100000, **kw))
        err = max(float(np.abs(got[c * nb + k] - cpu.blocks[k].points).max()) for k in range(nb))
        chord = 0.0799 * sc
        tol = 1e-9 * chord   # 3 outer iterations: the mesh is still regular, fp64 reproduces the exact sequence (tests/golden/t106_white_truth.npz)
        assert err <= tol, (c, err, tol)


def test_every_cut_of_a_batch_is_solved_as_its_own_system(gpu_lib, orc, monkeypatch):
    """The reference meshes the cuts one after the other: each has its own ||b||, tolerance max(atol, rtol ||b||)
    (GMRES.zig:305-306 / BiCGStab.zig:291), iteration count and stopping test.  At the reference's default tolerances
    (rtol 1e-6) a batch must therefore behave cut by cut like separate runs -- round 1 tested one norm over the batch."""
    from turbomesh_b200 import smoothing

    base, z, meta = load_fixture("t106_white")
    scales = [1.0, 1.2, 0.25, 0.6]     # different ||b|| per cut (coordinates stay below 2: the TFI's re-computed interface nodes must agree within 1e-15)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver(method="picard_bicgstab", rtol=1e-6, atol=1e-8, max_inner_iterations=1000)
    nb = len(base.blocks)

    def run(sc, iterations):
        batch, groups = synthetic.batch_of_cuts(base, sc)
        with smoothing.DeviceMesh(batch, upload=False) as dm:
            for k, b in enumerate(batch.blocks):
                dm.tfi_block(k, *b.edge_args())
            dm.set_white_groups(groups)
            dm.begin_smoothing(sol, cf)
            st = dm.smooth(iterations, sol, cf)
            assert dm.component_count == len(sc)
            assert [dm.component_of_block(k) for k in range(len(batch.blocks))] == [k // nb for k in range(len(batch.blocks))]
            comps = [dm.component_stats(c) for c in range(len(sc))]
            return [dm.download_block(k) for k in range(len(batch.blocks))], st, comps

    got, st, comps = run(scales, 2)
    assert st["converged"] == 1
    for c, sc in enumerate(scales):
        rec = comps[c]
        assert rec["nodes"] == 25118 and rec["status"] == (1, 1)
        for xy in range(2):
            assert rec["tolerance"][xy] == max(1e-8, 1e-6 * rec["norm_b"][xy])
            assert rec["norm_r"][xy] <= rec["tolerance"][xy]            # every cut meets ITS tolerance
            assert rec["norm_b"][xy] == pytest.approx(comps[0]["norm_b"][xy] * sc / scales[0], rel=1e-9)
        alone, st1, (rec1,) = run([sc], 2)
        # the same cut on its own: the same system (||b||, tolerance), iteration counts of the same size (BiCGStab's path depends on
        # the rounding of differently ordered sums: +-20 %), the same mesh within the tolerance
        for xy in range(2):
            assert abs(rec["iterations"][xy] - rec1["iterations"][xy]) <= 0.25 * rec1["iterations"][xy] + 3, (c, rec, rec1)
            assert rec["tolerance"][xy] == pytest.approx(rec1["tolerance"][xy], rel=1e-12)
            assert rec["norm_b"][xy] == pytest.approx(rec1["norm_b"][xy], rel=1e-12)
        err = max(float(np.abs(got[c * nb + k] - alone[k]).max()) for k in range(nb))
        assert err <= 2e-3 * sc, (c, err)    # two solves that both stop at rtol 1e-6 (the exact Picard step is 7e-4 chord away from either)
    assert st["inner_iterations"] >= sum(sum(r["iterations"]) for r in comps) > 0    # the stats sum over both outer iterations, the records hold the last
    # independence, bit for bit: in the phased form (one launch per phase over all systems, the path batches take) a cut's solve
    # does not depend on which other cuts share the launches
    monkeypatch.setenv("TM_KRYLOV", "phased")
    a, _, ca = run(scales, 2)
    b, _, cb = run([scales[0], scales[3]], 2)
    for k in range(nb):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[3 * nb + k], b[nb + k])
    assert ca[0] == cb[0] and ca[3] == cb[1]
    # ||b|| is the norm of the reference's right-hand side (fixed rows carry their coordinates, ...): oracle's assembled rhs of cut 0
    single, _ = synthetic.batch_of_cuts(base, [scales[0]])
    cpu = synthetic.materialize(single, orc.tfi)
    O = orc.System(cpu, orc.options(control_function="white", ds_target=meta["ds_target"], theta_target=meta["theta_target"]))
    O.iterate(0)
    O.fill(1)
    for xy, y_mode in enumerate((False, True)):
        O.fill_specific(y_mode)
        rhs = O.csr()[3 + xy]
        assert comps[0]["norm_b"][xy] == pytest.approx(float(np.linalg.norm(rhs)), rel=1e-5)
    O.close()


@pytest.mark.parametrize("tile_rows", [8, 16])
def test_two_level_preconditioner_in_the_batch_path(gpu_lib, monkeypatch, tile_rows):
    """The coarse space of the phased launches (krylov_phased.cuh, on by default for batches with tall tiles): the cuts of a
    batch reach the same exact Picard sequence -- cut 0 (the unscaled T106) against the extended-precision truth, the others
    against the point-Jacobi run -- with well under 60 % of the Krylov iterations."""
    from turbomesh_b200 import smoothing
    from util import GOLDEN, chord_of

    base, z, meta = load_fixture("t106_white")
    tz = np.load(os.path.join(GOLDEN, "t106_white_truth.npz"))
    scales = [1.0, 1.2, 0.6]
    nb = len(base.blocks)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver.tight()
    monkeypatch.setenv("TM_KRYLOV", "phased")
    monkeypatch.setenv("TM_KRYLOV_TILE_ROWS", str(tile_rows))     # tall tiles: what a large batch gets (16 rows from 2 M nodes on; a small batch has 2-row tiles and no coarse space)
    res, its = {}, {}
    for coarse in ("0", "1"):
        monkeypatch.setenv("TM_KRYLOV_COARSE", coarse)
        batch, groups = synthetic.batch_of_cuts(base, scales)
        with smoothing.DeviceMesh(batch, upload=False) as dm:
            for k, b in enumerate(batch.blocks):
                dm.tfi_block(k, *b.edge_args())
            dm.set_white_groups(groups)
            dm.begin_smoothing(sol, cf)
            st = dm.smooth(8, sol, cf)
            res[coarse] = [dm.download_block(k) for k in range(len(batch.blocks))]
        assert st["converged"] == 1 and st["last_inner_residual"] <= 1e-13
        its[coarse] = st["inner_iterations"]
    chord = float(np.ptp(res["0"][0][:, 0, 0]))
    for coarse in ("0", "1"):
        err = max(float(np.abs(res[coarse][k] - tz[f"truth8_b{k}"]).max()) for k in range(nb))
        assert err <= 1e-9 * chord, (coarse, err / chord)
    for c, sc in enumerate(scales):
        d = max(float(np.abs(res["0"][c * nb + k] - res["1"][c * nb + k]).max()) for k in range(nb))
        assert d <= 1e-9 * chord * sc, (c, d / (chord * sc))
    assert its["1"] < 0.6 * its["0"], its
