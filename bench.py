#!/usr/bin/env python
"""bench.py -- node-updates/s of the TFI + elliptic-smoothing hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)

A *step* is one pass of the hot path over one synthetic mesh: TFI of every block from its (device-resident) edges,
then `--sweeps` smoothing sweeps of the whole mesh (interior rows, interface/junction/sliding rows, residual
reduction).  node-updates = nodes x sweeps.  Workload at EVERY N (so that the driver's scaling efficiency compares like
with like): BASELINE.json config 4, one block column of 8 blocks of 4097 x 2049 per GPU (64 blocks / 512 Mi nodes at
N = 8), weak scaling; `--workload single` is config 3 (one 8192^2 block), `--workload cuts` config 5.  After the timed
region every rank checks its blocks against the oracle's assembled rows (`parity_check`), the N = 1 line also carries the
reference's own configurations (`configs`: T106, LS89 x4, 128 cuts, the 8192^2 block) with the CPU port timed on the same
mesh in the same run.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "node-updates/s"
UNIT = "node-updates/s"
BYTES_PER_NODE_UPDATE = 32.0  # read own x,y (16 B) + write new x,y (16 B), Laplace control function
BYTES_PER_NODE_UPDATE_PQ = 48.0  # ... + 16 B of P,Q when the White control function is resident


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    summaries (profiles/ncu_traffic.json names the file each number was read from); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f).get(key)
        return (float(rec["dram_bytes_per_launch"]), rec["source"]) if rec else (None, None)
    except Exception:
        return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([p.strip() for p in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def pin_to_gpu_numa_node(index: int):
    """Runs this process on the CPUs NVML reports as closest to the GPU, so that pinned host buffers are allocated on the
    GPU's NUMA node (host<->device copies of the e2e leg otherwise cross the socket interconnect)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------------------------
# Workloads
# --------------------------------------------------------------------------------------------------------------
CASCADE_SAMPLE_BLOCK = {1: (257, 129), 2: (193, 97), 4: (129, 65), 8: (97, 49)}   # bounded CPU samples: ~3e5 nodes at every N


def cascade_spec(world, args, block=None):
    """Config 4 in tiling form: one block column (8 blocks) per GPU; the passage grows with N (length = N/8) so that the cells
    stay square, its waviness scales alike."""
    from turbomesh_b200 import synthetic

    ni, nj = block or (args.block_ni, args.block_nj)
    n_bj = args.blocks_per_gpu
    spec = synthetic.cascade(world, n_bj, ni, nj, length=world / 8.0, ay=0.015 * world / 8.0)
    owner = [bi for bi in range(world) for _ in range(n_bj)]
    return spec, owner


def passages_spec(world, args, factor=None):
    """Config 4 as named: `world` pitch-wise O4H passages around the T106 profile (the blade edges of the committed T106 fixture,
    moved to x >= 0), one passage per GPU, cell counts = T106.json's times `--passage-factor`."""
    spec0, z, meta = _fixture("t106_white")
    from inputgen import passages
    up, down = z["b0_x_i_min"].copy(), z["b1_x_i_min"].copy()
    x0 = min(up[:, 0].min(), down[:, 0].min())
    up[:, 0] -= x0
    down[:, 0] -= x0
    return passages.o4h_passages(up, down, meta["pitch"], n_passages=world, factor=factor or args.passage_factor)


def workload_name(kind, world, args):
    if kind == "passages":
        return (f"o4h_passages_{world}x8_blocks_factor_{args.passage_factor} (config 4: synthetic multi-block T106-topology cascade, {world} pitch-wise O4H "
                f"passages of 8 blocks, cell counts of examples/T106 x {args.passage_factor}, one passage per GPU, interface halo exchange once per sweep)")
    if kind == "single":
        return f"single_block_{args.size}x{args.size} (config 3: synthetic single-block fp64 grid, TFI + elliptic smoothing)"
    return (f"cascade_{world}x{args.blocks_per_gpu}_blocks_of_{args.block_ni}x{args.block_nj} (config 4: synthetic multi-block cascade passage, "
            f"{args.blocks_per_gpu} blocks per GPU, interface halo exchange once per sweep)")


# --------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm).  Only the functions of this section touch oracle/ as the thing timed.
# --------------------------------------------------------------------------------------------------------------
def cpu_sample(kind, world, args):
    """Reference algorithm on one host core on a bounded sample of the arm's workload (same topology, smaller blocks): TFI + one
    outer iteration with the reference's default solver (gmres + ilu0, rtol 1e-6; examples/T106/T106.json:29-33)."""
    from oracle import oracle as orc
    from turbomesh_b200 import synthetic

    if kind == "single":
        n = args.ref_size
        spec, what = synthetic.single_block(n, n), f"single block {n}x{n} (same analytic edges as the GPU workload)"
    elif kind == "passages":
        f = {1: 2}.get(world, 1)
        spec, _ = passages_spec(world, args, f)
        what = f"{world} O4H passages at cell-count factor {f} (the GPU workload's topology at reduced block size)"
    else:
        blk = CASCADE_SAMPLE_BLOCK.get(world, (97, 49))
        spec, _ = cascade_spec(world, args, blk)
        what = f"cascade of {world}x{args.blocks_per_gpu} blocks of {blk[0]}x{blk[1]} (the GPU workload's topology at reduced block size)"
    nodes = sum(b.size[0] * b.size[1] for b in spec.blocks)
    t0 = time.perf_counter()
    mesh = synthetic.materialize(spec, orc.tfi)
    t_tfi = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = orc.smooth_mesh(mesh, 1, orc.options())
    t_smooth = time.perf_counter() - t0
    updates = float(nodes) * st["matvecs"]  # one node-update = one application of the 9-point operator to one node
    return {"value": updates / (t_tfi + t_smooth), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{what}, {nodes} nodes: TFI + 1 outer iteration gmres/ilu0 rtol 1e-6, {st['matvecs']} operator applications, "
                      f"{t_tfi + t_smooth:.2f} s; the reference is single-threaded",
            "seconds": t_tfi + t_smooth, "tfi_seconds": t_tfi, "krylov_iterations": st["krylov_iterations"]}


def _fixture(name):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import load_fixture

    return load_fixture(name)


def cpu_reference_config(name, iterations=None):
    """The reference's own configuration `name` (tests/golden fixture: the block edges the O4H template produced from the
    reference's example files) through the oracle port with the reference's settings: TFI + `iterations` outer iterations,
    White control function, gmres + ilu0, rtol 1e-6 / atol 1e-8 / 1000 inner iterations, one host core."""
    from oracle import oracle as orc
    from turbomesh_b200 import synthetic

    spec, z, meta = _fixture(name)
    its = iterations or meta["iterations"]
    t0 = time.perf_counter()
    mesh = synthetic.materialize(spec, orc.tfi)
    t_tfi = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = orc.smooth_mesh(mesh, its, orc.options(control_function="white", ds_target=meta["ds_target"], theta_target=meta["theta_target"]))
    t_smooth = time.perf_counter() - t0
    nodes = mesh.num_nodes()
    return {"seconds": t_tfi + t_smooth, "tfi_seconds": t_tfi, "outer_iterations": its, "krylov_iterations": st["krylov_iterations"],
            "operator_applications": st["matvecs"], "node_updates_per_s": nodes * st["matvecs"] / (t_tfi + t_smooth), "not_converged_solves": st["not_converged"],
            "cores": 1, "kind": "port", "solver": "gmres+ilu0 rtol 1e-6 atol 1e-8 max 1000 (reference defaults)"}, mesh


def reference_configs_cpu(args):
    """CPU legs of the `configs` block: config 1 in full, config 2 for `--ref-ls89-iterations` of its 10 outer iterations
    (95 s in full), config 5 as one cut (the reference meshes cuts one after the other: per-cut time x cuts)."""
    out = {}
    out["config1_t106"], _ = cpu_reference_config("t106_white")
    out["config2_ls89_x4"], _ = cpu_reference_config("ls89x4_white", args.ref_ls89_iterations)
    out["config5_cut"] = dict(out["config1_t106"], note="one T106 cut; a batch of n cuts costs n times this on the reference (sequential, single-threaded)")
    return out


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return  # the reference is a single-process CPU program; other ranks exit without work
    kind = {"tiling": "cascade", None: "passages"}.get(args.workload, args.workload)
    vals = []
    for k in range(args.warmup + args.steps):
        smp = cpu_sample(kind, world, args)
        if k >= args.warmup:
            vals.append(smp)
    secs = float(np.mean([v["seconds"] for v in vals]))
    value = float(np.mean([v["value"] for v in vals]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(kind, world, args), "sample": vals[-1]["sample"],
                       "solver": "gmres+ilu0 rtol 1e-6 (reference defaults), 1 outer iteration per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_configs and world == 1:
        line["configs"] = reference_configs_cpu(args)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# Parity inside the bench (the oracle as checker): after the timed region one more sweep is taken apart on windows
# --------------------------------------------------------------------------------------------------------------
def parity_check(dm, spec, owner, rank, world, args, dist, kind):
    """One extra damped-Jacobi sweep of the timed mesh, checked against the oracle's assembled rows on windows cut out of the
    full-size blocks: interior rows, up to three ordinary connections (sub-ranges and reversed traversal as they come), a
    periodic pair and (N > 1) a connection whose two sides live on different GPUs; `connected` copies must be exact copies
    (periodic ones: equal up to the shift), on the same GPU and across GPUs.  Every rank checks its own blocks; rank 0 reports
    the worst case.  The run fails if anything is off: a fast wrong answer is not a bench value."""
    from oracle import rowcheck
    from turbomesh_b200 import smoothing

    one = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=args.omega, device=int(os.environ.get("LOCAL_RANK", "0")))
    sizes = [b.size for b in spec.blocks]
    n, depth = 64, 24

    def usable(c):
        return all(min(sizes[r.block]) >= depth + 2 for r in c.ranges) and abs(c.ranges[0].end - c.ranges[0].start) + 1 >= n + 6

    conns = [c for c in spec.connections if usable(c)]
    local = [c for c in conns if owner[c.ranges[0].block] == rank and owner[c.ranges[1].block] == rank]
    plain = [c for c in local if c.periodicity is None]
    chosen = [plain[k] for k in sorted({0, len(plain) // 2, len(plain) - 1})] if plain else []
    chosen += [c for c in local if c.periodicity is not None][:1]
    cross = {}   # rank -> its first connection whose side 1 lives on another rank (every rank computes the same table)
    for c in conns:
        r0 = owner[c.ranges[0].block]
        if owner[c.ranges[1].block] != r0 and r0 not in cross:
            cross[r0] = c
    mine = [b for b in range(len(spec.blocks)) if owner[b] == rank]
    need = {max(mine, key=lambda b: sizes[b][0] * sizes[b][1])}
    for c in chosen:
        need |= {c.ranges[0].block, c.ranges[1].block}
    for r, c in cross.items():
        if r == rank:
            need.add(c.ranges[0].block)
        if owner[c.ranges[1].block] == rank:
            need.add(c.ranges[1].block)
    before = {b: dm.download_block(b) for b in sorted(need)}
    st = dm.smooth(1, one)
    after = {b: dm.download_block(b) for b in sorted(need)}
    worst, windows, copies_exact, periodic_dev = 0.0, 0, True, 0.0
    big = max(mine, key=lambda b: sizes[b][0] * sizes[b][1])
    ni, nj = sizes[big]
    for i0 in sorted({1, ni // 2, ni - n - 1}):
        j0 = max(1, min(nj - n - 1, nj // 2))
        win = before[big][i0:i0 + n, j0:j0 + n]
        want = rowcheck.interior_update(win, args.omega)
        worst = max(worst, float(np.abs(after[big][i0:i0 + n, j0:j0 + n][1:-1, 1:-1] - want[1:-1, 1:-1]).max())); windows += 1

    def window(c, side):
        r = c.ranges[side]
        k0 = (abs(r.end - r.start) + 1 - n) // 2
        return k0

    def check(c, win_b, mini_b, line_b_after):
        """side-0 rows of connection c against the oracle; copies on side 1 against side 0"""
        nonlocal worst, windows, copies_exact, periodic_dev
        r0, r1 = c.ranges
        k0 = window(c, 0)
        wa, ma, ia = rowcheck.cut_window(before[r0.block], r0, k0, n, depth)
        want = rowcheck.connection_update(wa, r0.side, ma, win_b, r1.side, mini_b, args.omega, c.periodicity)
        got = np.array([after[r0.block][ia(k)] for k in range(n)])
        worst = max(worst, float(np.abs(got[1:-1] - want).max())); windows += 1
        if c.periodicity is None:
            copies_exact = copies_exact and bool(np.array_equal(got[1:-1], line_b_after[1:-1]))
        else:
            periodic_dev = max(periodic_dev, float(np.abs(got[1:-1] + np.array(c.periodicity) - line_b_after[1:-1]).max()))

    def side1(c, blocks_before, blocks_after):
        r1 = c.ranges[1]
        k0 = window(c, 0)
        wb, mb, ib = rowcheck.cut_window(blocks_before[r1.block], r1, k0, n, depth)
        return wb, mb, np.array([blocks_after[r1.block][ib(k)] for k in range(n)])

    for c in chosen:
        check(c, *side1(c, before, after))
    cross_mismatch = None
    if world > 1:
        payload = {r: side1(c, before, after) for r, c in cross.items() if owner[c.ranges[1].block] == rank}
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
        if rank in cross:
            c = cross[rank]
            src = gathered[owner[c.ranges[1].block]][rank]
            w0 = worst
            worst = 0.0
            check(c, *src)
            cross_mismatch = worst
            worst = max(worst, w0)
    rec = {"max_row_update_mismatch": worst, "windows": windows, "copies_exact": copies_exact, "periodic_copy_max_deviation": periodic_dev,
           "cross_gpu_interface_mismatch": cross_mismatch, "sweep_max_update": st["last_max_update"]}
    if world > 1:
        allrec = [None] * world
        dist.all_gather_object(allrec, rec)
        cm = [r["cross_gpu_interface_mismatch"] for r in allrec if r["cross_gpu_interface_mismatch"] is not None]
        rec = {"max_row_update_mismatch": max(r["max_row_update_mismatch"] for r in allrec), "windows": sum(r["windows"] for r in allrec),
               "copies_exact": all(r["copies_exact"] for r in allrec), "periodic_copy_max_deviation": max(r["periodic_copy_max_deviation"] for r in allrec),
               "cross_gpu_interface_mismatch": max(cm) if cm else None, "cross_gpu_connections_checked": len(cm),
               "sweep_max_update": st["last_max_update"], "ranks_checked": world}
    rec["tolerance"] = 5e-14
    rec["what"] = ("one extra sweep after the timed region vs the damped-Jacobi update computed from the oracle's assembled CSR rows (interior, interface, "
                   "periodic and cross-GPU interface windows of the full-size blocks); copies compared bit for bit")
    ok = rec["max_row_update_mismatch"] <= rec["tolerance"] and rec["copies_exact"] and rec["periodic_copy_max_deviation"] <= 1e-15
    rec["ok"] = bool(ok)
    return rec


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def time_to_converged(dm, my_blocks, stream, torch, dist, world, local, nodes_total, repeats):
    """TFI + FAS multigrid V(3,3) until the mesh changes by <= 1e-10 (chord / passage height are O(1)) per cycle."""
    from turbomesh_b200 import smoothing

    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-10, device=local)
    best, cold = None, None
    for _ in range(repeats):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        evs[0].record(stream)
        for b in my_blocks:
            dm.tfi_block_resident(b)
        dm.begin_smoothing(mg)
        st_mg = dm.smooth(100, mg)
        evs[1].record(stream)
        dm.synchronize()
        tt = torch.tensor([evs[0].elapsed_time(evs[1]) * 1e-3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_all = float(tt.item())
        if cold is None:
            cold = t_all   # the first run also builds the multigrid hierarchy (one-off per topology)
        if best is None or t_all < best[0]:
            best = (t_all, st_mg)
    ops = best[1]["operator_applications"]
    # what "converged" means is checked independently of the cycle's own measure: one plain Jacobi sweep (omega 1) of the
    # converged mesh must move no node by more than ~1e-10 either (a diverged, non-finite mesh fails this outright)
    probe = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=1.0, device=local)
    pst = dm.smooth(1, probe)
    resid, sums = pst["last_max_update"], pst["last_sumsq_x"] + pst["last_sumsq_y"]
    finite = bool(np.isfinite(dm.download_block(my_blocks[0])).all()) and sums == sums   # (a max-norm ignores NaNs, the sums do not)
    if not (finite and resid <= 1e-8):
        raise SystemExit(f"time_to_converged: the multigrid result is not a fixed point of the sweep (Jacobi update {resid}, finite {finite})")
    return {"seconds": best[0], "solver_seconds": best[1]["gpu_seconds"], "cycles": best[1]["outer_iterations"], "jacobi_update_of_converged_mesh": resid,
            "criterion": "max-norm change of the mesh over one V(3,3) cycle <= 1e-10 (chord / passage height are O(1))", "last_max_update": best[1]["last_max_update"],
            "fine_grid_operator_applications": ops, "cold_seconds_incl_hierarchy_setup": cold,
            "solver": "TFI + geometric FAS multigrid over the whole block topology, damped-Jacobi smoother (omega 0.8), Anderson(3) on level-1 samples",
            "equivalent_node_updates_per_s": nodes_total * ops / best[1]["gpu_seconds"],
            "hbm_fraction_of_equivalent_sweeps": nodes_total * ops / best[1]["gpu_seconds"] * BYTES_PER_NODE_UPDATE / 1e9 / (measured_peak()[0] * world),
            "note": "best of %d; includes TFI and begin_smoothing; max over ranks" % repeats}


def passages_multigrid_status(dm, my_blocks, torch, dist, world, local, cycles=20):
    """The FAS multigrid on the O4H passage topology: it does NOT converge there yet -- a bounded attempt
    (V(3,3), no Anderson step, `cycles` cycles from the TFI mesh) is reported as it is; the time to a converged mesh of config 4
    is measured on the tiling form of the same size (`time_to_converged` of the line)."""
    from turbomesh_b200 import smoothing

    mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, device=local)
    os.environ["TM_MG_AA"] = "0"
    try:
        for b in my_blocks:
            dm.tfi_block_resident(b)
        dm.begin_smoothing(mg)
        first = dm.smooth(1, mg)["last_max_update"]
        st = dm.smooth(cycles - 1, mg)
    finally:
        os.environ.pop("TM_MG_AA", None)
    finite = bool(np.isfinite(dm.download_block(my_blocks[0])).all())
    probe = smoothing.CudaSolver(method="relax", sweeps_per_iteration=1, omega=1.0, device=local)
    resid = dm.smooth(1, probe)["last_max_update"] if finite else None
    return {"seconds": None, "status": "not converged (stalls)" if finite else "not converged (diverges)", "cycles_run": cycles, "mesh_change_first_cycle": first, "mesh_change_last_cycle": st["last_max_update"],
            "finite": finite, "jacobi_update_after_the_cycles": resid, "solver_seconds": st["gpu_seconds"],
            "note": "the geometric FAS multigrid does not converge on the O4H passage topology yet (it stalls at a mesh change of ~1e-4 per cycle on "
                    "passages of <= 1.5 M nodes and diverges on finer ones, DESIGN.md section 4): the time to a converged mesh of config 4 is reported "
                    "on the tiling form (time_to_converged)"}


def measure_sweeps(args, kind, torch, dist, rank, world, local, barrier, steps, warmup, with_ttc, with_e2e, with_parity):
    """The timed region of one workload: K steps of (TFI of every block + begin_smoothing + `sweeps` sweeps)."""
    from turbomesh_b200 import smoothing, synthetic

    sweeps = args.sweeps
    solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=args.omega, device=local)
    stream = torch.cuda.Stream()                         # the library launches on this stream, so torch events see its kernels
    if kind == "single":
        spec = synthetic.single_block(args.size, args.size)
        owner = [0]
        dm = smoothing.DeviceMesh(spec, device=local, stream=stream.cuda_stream, upload=False)
    else:
        spec, owner = passages_spec(world, args) if kind == "passages" else cascade_spec(world, args)
        uid = [smoothing.dist_unique_id() if (rank == 0 and world > 1) else None]
        if world > 1:
            dist.broadcast_object_list(uid, src=0)
        dm = smoothing.DeviceMesh(spec, device=local, stream=stream.cuda_stream, upload=False, owner=owner, rank=rank, n_ranks=world, unique_id=uid[0])
    my_blocks = [b for b in range(len(spec.blocks)) if owner[b] == rank]
    nodes_local = sum(spec.blocks[b].size[0] * spec.blocks[b].size[1] for b in my_blocks)
    nodes_total = sum(b.size[0] * b.size[1] for b in spec.blocks)
    for b in my_blocks:  # upload the edges once: afterwards the TFI inputs are resident in HBM
        dm.tfi_block(b, *spec.blocks[b].edge_args())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        for b in my_blocks:
            dm.tfi_block_resident(b)
        dm.begin_smoothing(solver)
        return dm.smooth(1, solver)

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = smoothing.kernel_launch_count()
    sweep_seconds, sweep_launches = 0.0, 0
    t0 = time.perf_counter()
    stats = None
    ev0.record(stream)
    for _ in range(steps):
        stats = step()
        sweep_seconds += stats["gpu_seconds"]          # CUDA events on the library's stream around the sweep loop only
        sweep_launches += sweeps
    ev1.record(stream)
    dm.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    elapsed = ev0.elapsed_time(ev1) * 1e-3              # device time of exactly K steps on the launching stream
    launches = smoothing.kernel_launch_count() - launches0
    clocks = sampler.stop()
    t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")   # max over ranks
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    value = nodes_total * sweeps * steps / elapsed
    parity = parity_check(dm, spec, owner, rank, world, args, dist, kind) if with_parity else None
    ttc = None
    if with_ttc and kind == "passages":
        ttc = passages_multigrid_status(dm, my_blocks, torch, dist, world, local)
    elif with_ttc:
        ttc = time_to_converged(dm, my_blocks, stream, torch, dist, world, local, nodes_total, 3 if kind == "single" else 2)
    e2e = None
    if with_e2e and kind == "single":
        e2e = run_e2e(args, spec, solver, torch)
    elif with_e2e:
        e2e = run_e2e_cascade(args, spec, dm, my_blocks, solver, torch, dist, world, nodes_total, barrier)
    peak, peak_src = measured_peak()
    per_launch = sweep_seconds / max(sweep_launches, 1)
    achieved = BYTES_PER_NODE_UPDATE * nodes_local / per_launch / 1e9
    traffic, traffic_src = ncu_traffic("single_block_8192x8192" if (kind == "single" and args.size == 8192) else
                                       (f"o4h_passage_factor_{args.passage_factor}" if kind == "passages" else f"cascade_column_{args.blocks_per_gpu}x{args.block_ni}x{args.block_nj}"))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "winslow_interior_bulk_kernel<RELAX>", "algorithmic_bytes_per_launch": BYTES_PER_NODE_UPDATE * nodes_local,
                "avg_launch_ms": per_launch * 1e3, "peak_source": peak_src + ", sustained copy figure"}
    out = {"value": value, "ms_per_step": elapsed / steps * 1e3, "nodes_total": nodes_total, "nodes_local": nodes_local, "roofline": roofline, "clocks": clocks,
           "gpu_launches": launches, "wall_ms_per_step": wall / steps * 1e3, "last_max_update": stats["last_max_update"] if stats else None,
           "halo_exchange": dm.halo_path, "parity_check": parity, "time_to_converged": ttc, "e2e": e2e}
    dm.close()
    return out


def gpu_reference_config(name, torch, local, iterations=None, repeats=3):
    """Configs 1 / 2 through the CUDA path with the reference's settings (White, rtol 1e-6 / atol 1e-8 / 1000, 10 outer
    iterations): device-resident (TFI from resident edges + smoothing, CUDA events) and end to end through tm_tfi_block +
    tm_smooth_mesh with host buffers (wall clock)."""
    from turbomesh_b200 import smoothing, synthetic

    spec, z, meta = _fixture(name)
    its = iterations or meta["iterations"]
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver(method="picard_bicgstab", rtol=1e-6, atol=1e-8, max_inner_iterations=1000, device=local)
    stream = torch.cuda.Stream()
    best = None
    with smoothing.DeviceMesh(spec, device=local, stream=stream.cuda_stream, upload=False) as dm:
        for k, b in enumerate(spec.blocks):
            dm.tfi_block(k, *b.edge_args())
        for _ in range(repeats + 1):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record(stream)
            for k in range(len(spec.blocks)):
                dm.tfi_block_resident(k)
            dm.begin_smoothing(sol, cf)
            st = dm.smooth(its, sol, cf)
            ev[1].record(stream)
            dm.synchronize()
            sec = ev[0].elapsed_time(ev[1]) * 1e-3
            if best is None or sec < best[0]:
                best = (sec, st)
    sec, st = best
    nodes = st["nodes"]
    e2e_best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        mesh = synthetic.materialize(spec, smoothing.tfi_block)      # Block2d.init per block: edges H2D, TFI, block D2H
        smoothing.smooth_mesh(mesh, its, sol, cf)                    # smooth.mesh: blocks H2D, solve, blocks D2H
        dt = time.perf_counter() - t0
        e2e_best = dt if e2e_best is None else min(e2e_best, dt)
    peak = measured_peak()[0]
    rate = nodes * st["operator_applications"] / st["gpu_seconds"]
    return {"nodes": nodes, "outer_iterations": its, "seconds": sec, "solver_seconds": st["gpu_seconds"], "e2e_seconds_host_buffers": e2e_best,
            "krylov_iterations": st["inner_iterations"], "operator_applications": st["operator_applications"], "converged": st["converged"],
            "node_updates_per_s": rate, "roofline_frac_48B": rate * BYTES_PER_NODE_UPDATE_PQ / 1e9 / peak,
            "solver": "matrix-free BiCGStab (point-Jacobi) on the row-scaled system, one persistent cooperative kernel per outer iteration, two fused phases per "
                      "iteration; rtol 1e-6 atol 1e-8 max 1000"}


def gpu_cuts(args, torch, local, n_cuts, repeats=2):
    """Config 5 on one GPU: a batch of independent T106 cuts (replicas only), each solved as its own system."""
    from turbomesh_b200 import smoothing, synthetic

    base, z, meta = _fixture("t106_white")
    scales = [1.0 + 0.2 * k / max(n_cuts - 1, 1) for k in range(n_cuts)]
    batch, groups = synthetic.batch_of_cuts(base, scales)
    stream = torch.cuda.Stream()
    t0 = time.perf_counter()
    dm = smoothing.DeviceMesh(batch, device=local, stream=stream.cuda_stream, upload=False)
    t_create = time.perf_counter() - t0
    for k, b in enumerate(batch.blocks):
        dm.tfi_block(k, *b.edge_args())
    dm.set_white_groups(groups)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver(method="picard_bicgstab", rtol=1e-6, atol=1e-8, max_inner_iterations=1000, device=local)
    best = None
    for _ in range(repeats + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for k in range(len(batch.blocks)):
            dm.tfi_block_resident(k)
        dm.begin_smoothing(sol, cf)
        st = dm.smooth(meta["iterations"], sol, cf)
        ev[1].record(stream)
        dm.synchronize()
        sec = ev[0].elapsed_time(ev[1]) * 1e-3
        if best is None or sec < best[0]:
            best = (sec, st)
    sec, st = best
    # the caller in front of the path on the device as well: O4H blocking of the same number of cuts (T106 cell counts) from the
    # blade points -- spline fit, edge discretisation, normal offset, combination in six launches (turbomesh_b200/blocking.py)
    from turbomesh_b200.blocking import Cells, Cut, O4HBatch
    from turbomesh_b200.clustering import Roberts
    cells = Cells(o_grid=40, middle_i=100, in_up_j=30, in_down_j=10, in_i=10, out_up_j=40, out_down_j=10, out_i=10, down_j=40, bulge=40, upstream_i=20, downstream_i=10)
    cuts = [Cut(z["b0_x_i_min"] * sc, z["b1_x_i_min"] * sc, float(meta["pitch"]) * sc) for sc in scales]
    t_blocking = None
    for _ in range(3):   # the first batch of a process also pays the first touch of its host buffers (0.2 s); a stream of batches does not
        t0 = time.perf_counter()
        blocked, _ = O4HBatch(cells, Roberts(0.5, 1.03), device=local).run(cuts)
        dt = time.perf_counter() - t0
        t_blocking = dt if t_blocking is None else min(t_blocking, dt)
    assert len(blocked.blocks) == len(batch.blocks)
    worst = max(max(dm.component_stats(c)["norm_r"][xy] / dm.component_stats(c)["tolerance"][xy] for xy in range(2)) for c in range(0, n_cuts, max(1, n_cuts // 16)))
    dm.close()
    peak = measured_peak()[0]
    rate = st["nodes"] * st["operator_applications"] / st["gpu_seconds"]
    return {"cuts": n_cuts, "nodes": st["nodes"], "seconds": sec, "solver_seconds": st["gpu_seconds"], "create_seconds": t_create, "blocking_seconds": t_blocking, "cuts_per_second": n_cuts / sec,
            "krylov_iterations": st["inner_iterations"], "operator_applications": st["operator_applications"], "converged": st["converged"],
            "worst_sampled_residual_over_own_tolerance": worst, "node_updates_per_s": rate, "roofline_frac_48B": rate * BYTES_PER_NODE_UPDATE_PQ / 1e9 / peak,
            "solver": "matrix-free BiCGStab on the row-scaled systems, one launch per phase over all cuts, two-level preconditioner (aggregates of 16 x 8 nodes, "
                      "direct coarse solve); rtol 1e-6 atol 1e-8 max 1000",
            "note": "every cut is its own linear system (own ||b||, tolerance, iteration count).  node_updates_per_s counts the operator applications actually "
                    "done: the coarse space cuts them by 2.7x (3131 -> ~1170 per node), so the rate and its roofline fraction fall while the time to "
                    "solution improves (0.73 -> 0.40 s); an application still moves ~200 B per node through HBM at ~3 TB/s (latency-limited launches)"}


def configs_block(args, torch, local):
    """The reference's own configurations in the driver-run line (N = 1): GPU seconds next to the CPU port on the same mesh in
    the same run."""
    cpu = {} if args.no_cpu_baseline else reference_configs_cpu(args)
    out = {}
    g1 = gpu_reference_config("t106_white", torch, local)
    out["config1_t106"] = {"gpu": g1, "cpu": cpu.get("config1_t106"), "same_config": True}
    g2 = gpu_reference_config("ls89x4_white", torch, local)
    g2s = gpu_reference_config("ls89x4_white", torch, local, iterations=args.ref_ls89_iterations, repeats=1)
    out["config2_ls89_x4"] = {"gpu": g2, "gpu_same_iterations_as_cpu": g2s, "cpu": cpu.get("config2_ls89_x4"), "same_config": True,
                              "note": f"the CPU port runs {args.ref_ls89_iterations} of the 10 outer iterations (95 s in full); the ratio uses the GPU run of the same iterations"}
    g5 = gpu_cuts(args, torch, local, args.cuts_per_gpu)
    out["config5_cuts"] = {"gpu": g5, "cpu_one_cut": cpu.get("config5_cut"), "same_config": True,
                           "note": "CPU time of the batch = cuts x one cut (independent meshes, sequential single-threaded reference)"}
    for key, g, c, scale in (("config1_t106", g1, cpu.get("config1_t106"), 1.0), ("config2_ls89_x4", g2s, cpu.get("config2_ls89_x4"), 1.0),
                             ("config5_cuts", g5, cpu.get("config5_cut"), float(args.cuts_per_gpu))):
        if c:
            out[key]["cpu_over_gpu_seconds"] = c["seconds"] * scale / g["seconds"]
            if "e2e_seconds_host_buffers" in g:
                out[key]["cpu_over_gpu_seconds_e2e"] = c["seconds"] * scale / g["e2e_seconds_host_buffers"]
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "cuts":
        return run_cuts(args, torch, dist, rank, world, local, barrier)
    kind = {"tiling": "cascade", None: "passages"}.get(args.workload, args.workload)
    if kind == "single" and world != 1:
        raise SystemExit("the single-block workload does not shard; use --workload cascade for N > 1")
    m = measure_sweeps(args, kind, torch, dist, rank, world, local, barrier, args.steps, args.warmup, not args.no_ttc, not args.no_e2e, not args.no_parity)
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(kind, world, args), "nodes": m["nodes_total"], "sweeps_per_step": args.sweeps, "omega": args.omega,
                       "step": "TFI of all blocks from device-resident edges + begin_smoothing + sweeps (damped Jacobi, coefficients from the current iterate)",
                       "cache": "inputs (2 ping-pong fields of 16 B x nodes_per_gpu) are far larger than the 126 MB L2", "nodes_per_gpu": m["nodes_local"],
                       "halo_exchange": m["halo_exchange"]},
            "roofline": m["roofline"], "clocks": m["clocks"], "gpu_launches": m["gpu_launches"], "wall_ms_per_step": m["wall_ms_per_step"],
            "last_max_update": m["last_max_update"]}
    if m["e2e"]:
        line["e2e"] = m["e2e"]
    if m["parity_check"]:
        line["parity_check"] = m["parity_check"]
    if kind == "passages" and not args.no_ttc:
        # time to a converged mesh of config 4: the tiling form (same node count class, interfaces / periodic pairs / junctions /
        # sliding rows / walls) -- the multigrid does not converge on the O4H passage topology yet, its status rides along
        t = measure_sweeps(args, "cascade", torch, dist, rank, world, local, barrier, 3, 3, True, False, not args.no_parity)
        line["time_to_converged"] = dict(t["time_to_converged"], workload=workload_name("cascade", world, args),
                                          sweep_rate_node_updates_per_s=t["value"], sweep_roofline_frac=t["roofline"]["frac"], parity_check=t["parity_check"])
        line["multigrid_on_passages"] = m["time_to_converged"]
    else:
        line["time_to_converged"] = m["time_to_converged"] or {"seconds": None, "note": "skipped (--no-ttc)"}
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_sample(kind, world, args).items() if k in ("value", "unit", "cores", "kind", "sample")}
        if not args.no_configs:
            cfg = configs_block(args, torch, local)
            if kind != "single":   # config 3, the 8192^2 block, next to the headline workload
                s3 = measure_sweeps(args, "single", torch, dist, rank, world, local, barrier, max(3, min(args.steps, 5)), 3, not args.no_ttc, not args.no_e2e, not args.no_parity)
                cfg["config3_single_block"] = {"workload": workload_name("single", 1, args), **{k: s3[k] for k in ("value", "ms_per_step", "roofline", "time_to_converged", "e2e", "parity_check")}}
            line["configs"] = cfg
    ok = True
    for pc in [m["parity_check"], line.get("time_to_converged", {}).get("parity_check")] + ([line["configs"]["config3_single_block"]["parity_check"]] if "configs" in line and "config3_single_block" in line["configs"] else []):
        ok = ok and (pc is None or pc["ok"])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("parity_check failed: the timed mesh does not satisfy the oracle's rows")


def run_cuts(args, torch, dist, rank, world, local, barrier):
    """Config 5: a batch of independent T106 cuts per GPU (replicas only, no collective): TFI of all blocks + the
    reference's smoothing settings (10 outer iterations, White control function, default tolerances) per step."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import load_fixture

    from turbomesh_b200 import smoothing, synthetic

    base, z, meta = load_fixture("t106_white")
    n_cuts = args.cuts_per_gpu
    total_cuts = n_cuts * world
    scales = [1.0 + 0.2 * (rank * n_cuts + k) / max(total_cuts - 1, 1) for k in range(n_cuts)]
    batch, groups = synthetic.batch_of_cuts(base, scales)
    stream = torch.cuda.Stream()
    t0 = time.perf_counter()
    dm = smoothing.DeviceMesh(batch, device=local, stream=stream.cuda_stream, upload=False)
    t_create = time.perf_counter() - t0
    for k, b in enumerate(batch.blocks):
        dm.tfi_block(k, *b.edge_args())
    dm.set_white_groups(groups)
    cf = smoothing.White(meta["ds_target"], meta["theta_target"])
    sol = smoothing.CudaSolver(method="picard_bicgstab", rtol=1e-6, atol=1e-8, max_inner_iterations=1000, device=local)

    def step():
        for k in range(len(batch.blocks)):
            dm.tfi_block_resident(k)
        dm.begin_smoothing(sol, cf)
        return dm.smooth(meta["iterations"], sol, cf)

    for _ in range(args.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = smoothing.kernel_launch_count()
    ev0.record(stream)
    ops = 0
    for _ in range(args.steps):
        st = step()
        ops += st["operator_applications"]
    ev1.record(stream)
    dm.synchronize()
    barrier()
    elapsed = ev0.elapsed_time(ev1) * 1e-3
    t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    nodes_local = dm.node_count
    line = {"metric": METRIC, "value": nodes_local * world * ops / elapsed, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"batch of {total_cuts} T106 cuts ({n_cuts} per GPU, 8 blocks / 25118 nodes each, scaled 1.0..1.2; config 5), no communication",
                       "step": "TFI of all blocks + 10 outer iterations, White control function, BiCGStab rtol 1e-6 (reference defaults)",
                       "nodes": nodes_local * world, "cuts_per_second": total_cuts * args.steps / elapsed, "create_seconds": t_create},
            "gpu_launches": smoothing.kernel_launch_count() - launches0, "converged": st["converged"], "last_inner_residual": st["last_inner_residual"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dm.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, spec, solver, torch):
    """Same step through tm_tfi_block + tm_smooth_mesh with pinned HOST buffers: H2D/D2H inside the timed region."""
    from turbomesh_b200 import smoothing
    from turbomesh_b200.discrete import Block2d, Mesh

    b = spec.blocks[0]
    ni, nj = b.size
    pinned = torch.empty((ni, nj, 2), dtype=torch.float64, pin_memory=True)
    host = pinned.numpy()
    edges = b.edge_args()
    mesh = Mesh([Block2d.__new__(Block2d)], ["block"], [], [])
    mesh.blocks[0].points = host
    h2d = sum(a.nbytes for a in edges) + host.nbytes      # TFI edges + the mesh going into smooth.mesh
    d2h = 2 * host.nbytes                                   # TFI result + smoothed mesh

    def step():
        smoothing.tfi_block(*edges, out=host)               # Block2d.init: edges H2D, TFI, block D2H
        return smoothing.smooth_mesh(mesh, 1, solver)       # smooth.mesh: block H2D, sweeps, block D2H (in place)

    for _ in range(max(1, min(args.warmup, 3))):
        step()
    torch.cuda.synchronize()
    # what the host link of this box delivers (explains the gap between `value` and `e2e`)
    dev = torch.empty_like(pinned, device="cuda")
    link = {}
    for name, (dst, src) in {"h2d_gbs": (dev, pinned), "d2h_gbs": (pinned, dev)}.items():
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        link[name] = host.nbytes / (time.perf_counter() - t0) / 1e9
    del dev
    steps = max(3, min(args.steps, 5))
    per_step = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st = step()
        torch.cuda.synchronize()
        per_step.append(time.perf_counter() - t0)
    dt = float(np.median(per_step))   # host-side jitter (page faults, clock ramps) is large on these boxes: median of the steps
    return {"value": ni * nj * solver.sweeps_per_iteration / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": dt * 1e3, "ms_all_steps": [round(t * 1e3, 2) for t in per_step], "statistic": "median over the steps",
            "api": "tm_tfi_block + tm_smooth_mesh (host buffers in pinned memory)", "steps": steps, "host_link": link,
            "last_max_update": st["last_max_update"], "streamed_chunks": st["streamed_chunks"],
            "note": "tm_smooth_mesh streams the block through the device in row chunks (upload, sweeps and download of successive "
                    "chunks overlap; bit-identical to the resident path)" if st["streamed_chunks"] else "resident"}


def run_e2e_cascade(args, spec, dm, my_blocks, solver, torch, dist, world, nodes_total, barrier):
    """The multi-block step through the public handle API with HOST buffers, one process per GPU: every step the rank hands
    the edges of its blocks to ``tm_mesh_tfi_block`` from host memory (H2D inside), smooths collectively and reads its
    blocks back into pinned host memory (``tm_mesh_download_block``, D2H inside).  Wall clock, max over ranks."""
    outs = {b: torch.empty((spec.blocks[b].size[0], spec.blocks[b].size[1], 2), dtype=torch.float64, pin_memory=True) for b in my_blocks}
    host = {b: outs[b].numpy() for b in my_blocks}
    def pinned(a):   # the step's inputs live in page-locked host memory, like its outputs
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        keep.append(t)
        return t.numpy()

    keep = []
    edges = {b: tuple(pinned(np.ascontiguousarray(a, dtype=np.float64)) for a in spec.blocks[b].edge_args()) for b in my_blocks}
    h2d = sum(a.nbytes for b in my_blocks for a in edges[b])
    d2h = sum(host[b].nbytes for b in my_blocks)

    def step():
        for b in my_blocks:
            dm.tfi_block(b, *edges[b])          # Block2d.init: edges H2D + TFI
        dm.begin_smoothing(solver)
        st = dm.smooth(1, solver)               # smooth.mesh on the device-resident blocks
        for b in my_blocks:
            dm.download_block(b, host[b])       # copy-back (smooth.zig:139-153) into host memory
        return st

    def step_async():                           # the same step with the copy-back started, not awaited: the next step's TFI and
        for b in my_blocks:                     # sweeps run while this step's blocks travel to the host (snapshot on the device)
            dm.tfi_block(b, *edges[b])
        dm.begin_smoothing(solver)
        st = dm.smooth(1, solver)
        for b in my_blocks:
            dm.download_block_async(b, host[b])
        return st

    step()
    barrier()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        st = step()
    barrier()
    t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_serial = float(t.item())
    # steady state of a stream of such steps: every step still uploads its edges and brings all its blocks to the host inside
    # the timed region (the last copy is awaited before the clock stops); only the waiting is overlapped
    check = host[my_blocks[0]].copy()
    step_async(); dm.download_wait()
    same = bool(np.array_equal(check, host[my_blocks[0]]))
    barrier()
    steps_p = max(4, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps_p):
        st = step_async()
    dm.download_wait()
    barrier()
    t = torch.tensor([(time.perf_counter() - t0) / steps_p], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    # what the host link allows: the blocks come down AFTER the sweeps (every block is coupled to its neighbours until the last
    # sweep), so a step cannot be shorter than the sweeps plus the read-back at the link rate measured here with all ranks copying
    dev = torch.empty(int(d2h), dtype=torch.uint8, device="cuda")
    pin = torch.empty(int(d2h), dtype=torch.uint8, pin_memory=True)
    pin.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter(); pin.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    link = d2h / (time.perf_counter() - t0) / 1e9
    lt = torch.tensor([link], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.MIN)
    link = float(lt.item())
    del dev, pin
    compute_ms = st["gpu_seconds"] * 1e3
    ceiling_ms = compute_ms + d2h / (link * 1e9) * 1e3
    return {"value": nodes_total * solver.sweeps_per_iteration / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": dt * 1e3, "api": "tm_mesh_tfi_block (host edges) + tm_mesh_begin_smoothing + tm_mesh_smooth + tm_mesh_download_block_async (pinned host blocks), "
                                            "tm_mesh_download_wait before the clock stops; per rank",
            "steps": steps_p, "last_max_update": st["last_max_update"],
            "note": "bytes per rank.  A stream of steps: the copy-back of step k (snapshot on the device, then D2H on a stream of its own) overlaps the "
                    "TFI and sweeps of step k+1; every step's H2D and D2H are inside the timed region.",
            "async_result_equals_blocking": same,
            "one_step_at_a_time": {"value": nodes_total * solver.sweeps_per_iteration / dt_serial, "ms_per_step": dt_serial * 1e3, "steps": steps,
                                   "api": "the same with the blocking tm_mesh_download_block"},
            "host_link": {"d2h_gbs_per_rank_all_ranks_copying": link, "sweeps_ms": compute_ms, "read_back_ms_at_link_rate": d2h / (link * 1e9) * 1e3,
                          "step_floor_ms_one_step_at_a_time": ceiling_ms, "e2e_ceiling_one_step_at_a_time": nodes_total * solver.sweeps_per_iteration / (ceiling_ms * 1e-3),
                          "step_floor_ms_stream_of_steps": max(compute_ms, d2h / (link * 1e9) * 1e3),
                          "e2e_ceiling_stream_of_steps": nodes_total * solver.sweeps_per_iteration / (max(compute_ms, d2h / (link * 1e9) * 1e3) * 1e-3),
                          "note": "one step at a time cannot be shorter than sweeps + read-back at the measured link rate (every block is coupled to its "
                                  "neighbours until the last sweep); a stream of steps is bound by the larger of the two"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="turbomesh_b200", choices=["turbomesh_b200", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="single-block edge length (N=1)")
    ap.add_argument("--workload", default=None, choices=["passages", "tiling", "cascade", "single", "cuts"],
                    help="default: passages (config 4 as named: one O4H passage per GPU) at every N; tiling / cascade = config 4 as a Cartesian tiling")
    ap.add_argument("--passage-factor", type=int, default=48, help="passages: cell counts of examples/T106 times this (48 -> 55 M nodes per passage)")
    ap.add_argument("--cuts-per-gpu", type=int, default=128, help="--workload cuts: T106 cuts per GPU (1024 cuts on 8 GPUs)")
    ap.add_argument("--block-ni", type=int, default=4097, help="cascade: nodes per block along i (2^k + 1 keeps every multigrid level nested)")
    ap.add_argument("--block-nj", type=int, default=2049)
    ap.add_argument("--blocks-per-gpu", type=int, default=8)
    ap.add_argument("--sweeps", type=int, default=100, help="smoothing sweeps per step")
    ap.add_argument("--omega", type=float, default=0.9)
    ap.add_argument("--ref-size", type=int, default=512, help="edge length of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ttc", action="store_true", help="skip the time-to-converged (multigrid) measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle-row check of the timed mesh")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (configs 1, 2, 3, 5 next to the headline workload; N = 1 only)")
    ap.add_argument("--ref-ls89-iterations", type=int, default=2, help="outer iterations of config 2 the CPU port runs (10 take ~95 s)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
