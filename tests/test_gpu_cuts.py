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
