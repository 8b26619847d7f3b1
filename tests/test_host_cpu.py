"""CPU tests of the host-side mirror (clustering, spline, Edge.combine, O4H, csv) against the reference's own
known-answer tests, and of the C-ABI surface."""
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from turbomesh_b200.clustering import Roberts, SingleHyperbolicClustering, Uniform
from inputgen.edges import EdgeView, combine
from turbomesh_b200.discrete import Edge
from inputgen.geometry import Line
from inputgen.spline import FittingSpline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _two_edges():
    e1 = Edge.init(3, Line((0.0, 0.0), (2.0, 0.0)), Uniform())
    e2 = Edge.init(3, Line((2.0, 0.0), (4.0, 0.0)), Uniform())
    return e1, e2


def test_edge_combine_known_answers():
    """The four cases of the reference's `combining edges` test (discrete.zig:219-290), exact equality."""
    e1, e2 = _two_edges()
    e = combine([EdgeView(e1, 0, 2), EdgeView(e2, 0, 2)])
    assert np.array_equal(e.points, [[0, 0], [1, 0], [2, 0], [3, 0], [4, 0]]) and np.array_equal(e.clustering, [0, 0.25, 0.5, 0.75, 1.0])
    e = combine([EdgeView(e1, 1, 2), EdgeView(e2, 0, 1)])
    assert np.array_equal(e.points, [[1, 0], [2, 0], [3, 0]]) and np.array_equal(e.clustering, [0, 0.5, 1.0])
    e = combine([EdgeView(e2, 2, 0), EdgeView(e1, 2, 0)])
    assert np.array_equal(e.points, [[4, 0], [3, 0], [2, 0], [1, 0], [0, 0]]) and np.array_equal(e.clustering, [0, 0.25, 0.5, 0.75, 1.0])
    e = combine([EdgeView(e2, 1, 0), EdgeView(e1, 2, 1)])
    assert np.array_equal(e.points, [[3, 0], [2, 0], [1, 0]]) and np.array_equal(e.clustering, [0, 0.5, 1.0])


def test_spline_straight_line_known_answer():
    """spline.zig:235-264."""
    pts = [(0.0, 0.0), (0.5, 0.5), (1.0, 1.0), (2.0, 2.0), (3.0, 3.0), (4.0, 4.0)]
    s = FittingSpline(pts, 3)
    v = s.interpolate([0.0, 0.125, 0.25, 0.5, 0.75, 1.0])
    assert np.abs(v - np.array(pts)).max() < 1e-9
    assert abs(s.integrate() - math.sqrt(2.0) * 4.0) < 1e-9


def test_spline_monotonic_and_two_point_length():
    """spline.zig:266-304."""
    pts = [(0.0, 0.0), (1.0, 0.5), (2.0, 1.5), (2.5, 3.0)]
    v = FittingSpline(pts, 3).interpolate([0.0, 0.5, 1.0])
    assert v[0, 0] <= v[1, 0] <= v[2, 0]
    assert np.abs(v[0] - pts[0]).max() < 1e-9 and np.abs(v[2] - pts[-1]).max() < 1e-9
    assert abs(FittingSpline([(0.0, 0.0), (0.0, 3.0)], 3).integrate() - 3.0) < 1e-9


def test_clusterings_hit_end_points_exactly():
    for f in (Uniform(), SingleHyperbolicClustering(0.01), SingleHyperbolicClustering(0.4 / 63)):
        u = f.compute(64)
        assert u[0] == 0.0 and u[-1] == 1.0 and np.all(np.diff(u) > 0)
    u = Roberts(0.5, 1.03).compute(41)
    assert abs(u[0]) < 1e-15 and abs(u[-1] - 1) < 1e-15 and np.all(np.diff(u) > 0)
    assert abs(u[20] - 0.5) < 1e-12  # alpha = 0.5 clusters symmetrically at both ends
    with pytest.raises(ValueError):
        SingleHyperbolicClustering(0.1).compute(41)  # (n-1) * delta_s > 1 (clustering.zig:68-76)


@pytest.mark.skipif(not os.path.exists(REF), reason="needs the reference's example data")
def test_csv_and_t106_profile_known_answers():
    """csv.zig:59-67 (first / last row) and spline.zig:306-514 (T106 surface length 0.4947 +- 1e-2, here the CSV profile
    is in metres with chord ~0.1 m, so only the CSV rows and the profile consistency are pinned)."""
    from inputgen.input import Input, parse_csv_into_vec2d

    d = parse_csv_into_vec2d(f"{REF}/examples/T106/T106_ps.dat")
    assert tuple(d[0]) == (1.127030384, -0.047185256) and tuple(d[-1]) == (1.047805900, 0.000076595)
    inp = Input.from_json(open(f"{REF}/examples/T106/T106.json").read())
    geom = inp.geometry(REF)
    assert geom.pitch == 0.08836
    assert 0.10 < geom.profile.down_part.total_length < 0.12 and 0.11 < geom.profile.up_part.total_length < 0.14


@pytest.mark.skipif(not os.path.exists(REF), reason="needs the reference's example data")
def test_o4h_template_reproduces_committed_fixture_inputs(orc):
    """The committed T106 fixture inputs are what the O4H mirror produces from the reference's example files."""
    from util import load_fixture
    from inputgen.input import Input

    inp = Input.from_json(open(f"{REF}/examples/T106/T106.json").read())
    calls = []

    def rec(*a):
        calls.append([np.array(x) for x in a])
        return orc.tfi(*a)

    mesh = inp.template.run(inp.geometry(REF), tfi=rec)
    spec, z, meta = load_fixture("t106_white")
    assert len(calls) == 8 and len(mesh.connections) == 21 and len(mesh.boundary_conditions) == 2
    assert sum(1 for c in mesh.connections if c.periodicity is not None) == 3
    for k, call in enumerate(calls):
        for got, want in zip(call, spec.blocks[k].edge_args()):
            assert np.array_equal(got, want)
    assert mesh.connections == spec.connections and mesh.boundary_conditions == spec.boundary_conditions


# ------------------------------------------------------------------------------------------ C ABI
def _declared_functions():
    text = open(os.path.join(ROOT, "include", "turbomesh_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(gpu_lib):
    """The C-ABI library loads (no GPU needed) and exports exactly what include/turbomesh_gpu.h declares."""
    from turbomesh_b200 import _lib

    declared = _declared_functions()
    assert len(declared) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("tm_"))
    assert exported == declared
    for name in declared:
        assert hasattr(gpu_lib, name)
    assert gpu_lib.tm_abi_version() == 1


def test_no_cpu_fallback_without_device(gpu_lib):
    """Without a CUDA device the product path fails loudly (TM_ERR_NO_DEVICE) instead of computing on the CPU."""
    import ctypes as C

    from turbomesh_b200 import _lib, smoothing, synthetic

    n = C.c_int(0)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    b = synthetic.single_block(9, 7).blocks[0]
    with pytest.raises(_lib.TurbomeshGpuError) as e:
        smoothing.tfi_block(*b.edge_args())
    assert e.value.code == _lib.TM_ERR_NO_DEVICE and "no CPU fallback" in e.value.message


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "turbomesh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "turbomesh_oracle" not in text, f
                assert "inputgen" not in text, f    # the INPUT-GEN restatements (tests/inputgen) are not product code either


def test_o4h_passages_rewire_the_pitchwise_periodic_connections(orc):
    """Config 4 as named (tests/inputgen/passages.py): k pitch-wise O4H passages; the template's three periodic connections
    become ordinary connections between neighbouring passages and stay periodic (k * pitch) between the last and the first.
    The oracle accepts the mesh (connectionDataCheck, topology rules) and classifies it like k copies of one passage."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from inputgen import passages
    from util import load_fixture

    from turbomesh_b200 import smoothing, synthetic

    spec, z, meta = load_fixture("t106_white")
    one = orc.System(synthetic.materialize(spec, orc.tfi))
    base_kinds = np.bincount(one.kinds(), minlength=5)
    n_junctions = len(one.junctions())
    one.close()
    for k in (1, 3):
        mesh, owner = passages.o4h_passages(z["b0_x_i_min"], z["b1_x_i_min"], meta["pitch"], n_passages=k, factor=1, o_grid_delta_s=0.01)
        assert len(mesh.blocks) == 8 * k and len(mesh.connections) == 21 * k and owner == [p for p in range(k) for _ in range(8)]
        per = [c for c in mesh.connections if c.periodicity is not None]
        assert len(per) == 3 and all(c.periodicity == (0.0, k * meta["pitch"]) for c in per)
        assert all(c.ranges[0].block <= c.ranges[1].block for c in mesh.connections)
        m = synthetic.materialize(mesh, orc.tfi)
        S = orc.System(m)
        assert np.array_equal(np.bincount(S.kinds(), minlength=5), k * base_kinds) and len(S.junctions()) == k * n_junctions
        S.close()
        if k == 1:   # one passage is the T106 example mesh: same block sizes; the O-grid blocks coincide up to the re-fit of the
            # profile splines through the fixture's edge nodes (the example's explicit inlet / outlet distances are not used)
            ref = synthetic.materialize(spec, orc.tfi).blocks
            assert [a.points.shape for a in m.blocks] == [b.points.shape for b in ref]
            assert max(np.abs(a.points - b.points).max() for a, b in zip(m.blocks[:2], ref[:2])) < 1e-5
    # the multigrid hierarchy of the refined passage: cell counts x 8 halve three times (and once more, like T106 itself)
    mesh, _ = passages.o4h_passages(z["b0_x_i_min"], z["b1_x_i_min"], meta["pitch"], n_passages=2, factor=8)
    assert len(smoothing.mg_plan(mesh)) >= 4


# ---- host-only planning of the multi-block multigrid hierarchy (tm_mg_plan) -----------------------------------------
def test_mg_plan_cascade_halves_every_block_down_to_3x3(gpu_lib):
    from turbomesh_b200 import smoothing, synthetic

    plan = smoothing.mg_plan(synthetic.cascade(2, 2, 65, 33))
    assert plan[0] == [(65, 33)] * 4
    assert plan[1] == [(33, 17)] * 4 and plan[2] == [(17, 9)] * 4
    assert plan[-1] == [(3, 3)] * 4
    # 33 nodes along j bottom out first (2 intervals); i goes on alone for one more level
    assert plan[-2] == [(5, 3)] * 4


def test_mg_plan_semi_coarsening_follows_the_cell_sizes(gpu_lib):
    from turbomesh_b200 import smoothing, synthetic

    mesh = synthetic.cascade(2, 2, 65, 65)
    iso = smoothing.mg_plan(mesh, cell_size=[1.0, 1.0] * 4)
    assert iso[1] == [(33, 33)] * 4
    thin_j = smoothing.mg_plan(mesh, cell_size=[1.0, 0.25] * 4)   # j spacing 4x finer: only j is halved until the cells are square
    assert thin_j[1] == [(65, 33)] * 4 and thin_j[2] == [(65, 17)] * 4 and thin_j[3] == [(33, 9)] * 4


def test_mg_plan_stops_at_odd_extents_and_odd_range_ends(gpu_lib):
    from turbomesh_b200 import smoothing, synthetic

    plan = smoothing.mg_plan(synthetic.cascade(2, 2, 12, 9))     # 11 x 8 intervals: i can never be halved
    assert [lv[0] for lv in plan] == [(12, 9), (12, 5), (12, 3)]
    t106, _, _ = __import__("util").load_fixture("t106_laplace")
    plan = smoothing.mg_plan(t106)
    assert len(plan) == 2                                         # 220 x 40 ... intervals halve once; then range ends become odd
    assert plan[1][0] == ((221 - 1) // 2 + 1, (41 - 1) // 2 + 1)


def test_mg_plan_single_block_without_connections(gpu_lib):
    from turbomesh_b200 import smoothing, synthetic

    plan = smoothing.mg_plan(synthetic.single_block(129, 65))
    assert [lv[0] for lv in plan][:3] == [(129, 65), (65, 33), (33, 17)] and plan[-1][0] == (3, 3)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) needs no GPU: one JSON line with the
    contract's keys."""
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-size", "40"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "node-updates/s" and line["unit"] == "node-updates/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_topology_analysis_of_a_batch_of_cuts_is_per_cut_and_fast(gpu_lib):
    """Independent cuts concatenated into one mesh: every cut contributes exactly the rows of a single cut (the junction
    discovery visits end-point pairs in the reference's order through an index, not by the O(n^2) scan), in linear time."""
    import time

    from turbomesh_b200 import smoothing, synthetic

    base, _, _ = __import__("util").load_fixture("t106_white")
    one = smoothing.dist_plan(base, [0] * len(base.blocks), 0, 1)
    n = 96
    batch, groups = synthetic.batch_of_cuts(base, [1.0 + 0.002 * k for k in range(n)])
    t0 = time.perf_counter()
    many = smoothing.dist_plan(batch, [0] * len(batch.blocks), 0, 1)
    assert time.perf_counter() - t0 < 20.0
    for key in ("n_own", "n_smoothed", "n_junction", "n_sliding", "n_slaves"):
        assert many[key] == n * one[key], key
    assert groups[5] == (5 * len(base.blocks), 5 * len(base.blocks) + 1)


def test_topology_validation_reports_what_the_reference_only_asserts(gpu_lib):
    """the reference's `std.debug.assert`s vanish in ReleaseFast; the C ABI validates the
    topology on the host (no GPU needed: tm_dist_plan runs the same analysis as tm_mesh_create) and reports a status code
    plus a message instead of aborting."""
    from dataclasses import replace

    from turbomesh_b200 import _lib, smoothing, synthetic
    from turbomesh_b200.boundary import Side

    def set_range(m, c, s, **kw):
        cn = m.connections[c]
        rs = list(cn.ranges)
        rs[s] = replace(rs[s], **kw)
        m.connections[c] = replace(cn, ranges=tuple(rs))

    def swap(m):
        m.connections[0] = replace(m.connections[0], ranges=m.connections[0].ranges[::-1])

    def bad_condition(m):
        bc = m.boundary_conditions[0]
        m.boundary_conditions[0] = replace(bc, range=replace(bc.range, block=77))

    cases = [
        (lambda m: set_range(m, 0, 1, block=9), _lib.TM_ERR_TOPOLOGY, "out of bounds"),
        (lambda m: set_range(m, 0, 0, end=999), _lib.TM_ERR_TOPOLOGY, "out of bounds"),
        (lambda m: set_range(m, 0, 1, end=7), _lib.TM_ERR_TOPOLOGY, "differ in length"),
        (lambda m: (set_range(m, 0, 0, end=3), set_range(m, 0, 1, end=3)), _lib.TM_ERR_UNSUPPORTED, "at least 6 nodes"),  # smooth.zig:631
        (swap, _lib.TM_ERR_TOPOLOGY, "smooth.zig:627"),
        (lambda m: set_range(m, 0, 1, block=0, side=Side.j_min), _lib.TM_ERR_TOPOLOGY, ""),  # same-block connections: i_min -> i_max only
        (bad_condition, _lib.TM_ERR_TOPOLOGY, "condition 0"),
    ]
    for mutate, code, text in cases:
        mesh = synthetic.cascade(2, 2, 12, 9)
        mutate(mesh)
        with pytest.raises(_lib.TurbomeshGpuError) as e:
            smoothing.dist_plan(mesh, [0] * 4, 0, 1)
        assert e.value.code == code and text in e.value.message, (e.value.code, e.value.message)
    with pytest.raises(_lib.TurbomeshGpuError, match="owner 5 out of range"):
        smoothing.dist_plan(synthetic.cascade(2, 2, 12, 9), [0, 0, 0, 5], 0, 2)


def test_stream_plan_covers_every_row_once_with_enough_margin(gpu_lib, monkeypatch):
    """Host-only plan of the streamed tm_smooth_mesh (streamed.inl): the chunks own every row exactly once, every window
    keeps `sweeps` rows between an artificial edge and the rows it owns, and reads owned rows of its direct neighbours
    only (the order of uploads and downloads relies on that)."""
    import random

    from turbomesh_b200 import smoothing

    assert smoothing.stream_plan(512, 512, 10) is None          # below TM_STREAM_MIN_NODES (default 8 Mi nodes): resident
    w, first, owned = smoothing.stream_plan(8192, 8192, 100)    # the bench's end-to-end step
    assert len(first) == 8 and w * 8 < 1.2 * 8192
    monkeypatch.setenv("TM_STREAM_MIN_NODES", "0")
    rng = random.Random(7)
    n_plans = 0
    for _ in range(3000):
        ni, t = rng.randint(3, 30000), rng.randint(1, 500)
        plan = smoothing.stream_plan(ni, 64, t)
        if plan is None:
            assert ni // (4 * t) < 2 or ni < 9
            continue
        n_plans += 1
        w, first, owned = plan
        k_n = len(first)
        assert 2 <= k_n <= 8 and owned[0] == 0 and owned[-1] == ni and first[0] == 0 and first[-1] + w == ni
        for k in range(k_n):
            assert owned[k] < owned[k + 1] and first[k] <= owned[k] and owned[k + 1] <= first[k] + w
            assert first[k] == 0 or owned[k] - first[k] >= t
            assert first[k] + w == ni or first[k] + w - owned[k + 1] >= t
            assert k < 2 or first[k] >= owned[k - 1]
    assert n_plans > 1000
    monkeypatch.setenv("TM_STREAM", "0")
    assert smoothing.stream_plan(8192, 8192, 100) is None


def test_spline_t106a_surface_length_known_answer():
    """spline.zig:306-514 ("T106 blade coordinate integration"): the spline through the published T106A contour (184 points
    over chord, chord 198 mm) is 264.7 mm + 230.0 mm long within 1e-2.  Table and expected value: tests/golden/
    t106a_blade_table.npz (written by tests/golden/make_spline_fixture.py from the reference's test)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "t106a_blade_table.npz"))
    spline = FittingSpline(z["points_over_chord"] * float(z["chord"]))
    assert abs(spline.integrate() - float(z["surface_length"])) <= float(z["tolerance"])
    # the length is the sum over the 200-sample arc table (spline.zig:74-110), slightly below the contour's polyline
    pts = z["points_over_chord"] * float(z["chord"])
    polyline = float(np.sqrt((np.diff(pts, axis=0) ** 2).sum(axis=1)).sum())
    assert 0.99 * polyline < spline.integrate() < polyline


def test_o4h_template_tables_match_the_sequential_restatement():
    """blocking.py keeps the O4H template as data (edge lengths, views of the combined edges, connections, conditions in terms
    of the cell counts); the sequential host restatement of O4H.run (tests/inputgen/templates.py) builds the same topology
    edge by edge -- both must agree for any cell counts."""
    from inputgen.geometry import Geometry, Profile
    from inputgen.templates import O4H, NumCells
    from turbomesh_b200.blocking import BLOCKS, Cells, _BLOCK_EDGES, _lengths, connections
    from turbomesh_b200.clustering import Uniform
    from util import load_fixture

    spec, z, meta = load_fixture("t106_white")
    up, down, pitch = z["b0_x_i_min"], z["b1_x_i_min"], float(meta["pitch"])
    for counts in (dict(o_grid=8, middle_i=20, in_up_j=6, in_down_j=4, in_i=5, out_up_j=8, out_down_j=4, out_i=5, down_j=6, bulge=6, upstream_i=4, downstream_i=3),
                   dict(o_grid=5, middle_i=9, in_up_j=3, in_down_j=7, in_i=4, out_up_j=5, out_down_j=6, out_i=3, down_j=8, bulge=4, upstream_i=6, downstream_i=7)):
        sizes = []
        one = O4H(blade_clustering=Uniform(), num_cells=NumCells(**counts)).run(
            Geometry(pitch, Profile(down, up)), tfi=lambda *a: sizes.append((len(a[4]), len(a[6]))) or np.zeros((len(a[4]), len(a[6]), 2)))
        n = _lengths(Cells(**counts))
        assert list(one.names) == list(BLOCKS)
        assert sizes == [(n[_BLOCK_EDGES[b][0]], n[_BLOCK_EDGES[b][2]]) for b in BLOCKS]
        conns, conds = connections(Cells(**counts), pitch)
        assert len(conns) == len(one.connections) == 21
        for a, b in zip(conns, one.connections):
            assert [(r.block, r.side, r.start, r.end) for r in a.ranges] == [(r.block, r.side, r.start, r.end) for r in b.ranges]
            assert a.periodicity == b.periodicity
        assert [(c.range.block, c.range.side, c.range.start, c.range.end, c.kind) for c in conds] == \
               [(c.range.block, c.range.side, c.range.start, c.range.end, c.kind) for c in one.boundary_conditions]
