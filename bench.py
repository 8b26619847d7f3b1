#!/usr/bin/env python
"""bench.py -- node-updates/s of the TFI + elliptic-smoothing hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)

A *step* is one pass of the hot path over one synthetic mesh: TFI of every block from its (device-resident) edges,
then `--sweeps` smoothing sweeps of the whole mesh (interior rows, interface/junction/sliding rows, residual
reduction).  node-updates = nodes x sweeps.  Workload: N=1 -> BASELINE.json config 3 (single block 8192 x 8192, the
largest single-GPU configuration); N>1 -> config 4 in tiling form, 8 blocks of 4097 x 2049 per GPU (64 blocks / 512 Mi
nodes at N=8), weak scaling.  Both also report the time to a converged mesh (TFI + FAS multigrid).  Prints ONE JSON line on rank 0.
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch of winslow_interior_kernel on the 8192^2 block, from the
# committed ncu capture (profiles/); None until measured.
NCU_TRAFFIC_BYTES_PER_LAUNCH = 2.150e9  # profiles/r1_ncu_winslow_interior_bulk_8192.txt: 1.1287 GB read + 1.0214 GB written


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
# CPU baseline (oracle port of the reference algorithm).  Only this function and run_reference() touch oracle/.
# --------------------------------------------------------------------------------------------------------------
def cpu_baseline_sample(n: int = 512):
    """Reference algorithm on one host core, bounded sample of the single-block workload: TFI + one outer iteration with
    the reference's default solver (gmres + ilu0, rtol 1e-6; examples/T106/T106.json:29-33)."""
    from oracle import oracle as orc
    from turbomesh_b200 import synthetic

    spec = synthetic.single_block(n, n)
    t0 = time.perf_counter()
    mesh = synthetic.materialize(spec, orc.tfi)
    t_tfi = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = orc.smooth_mesh(mesh, 1, orc.options())
    t_smooth = time.perf_counter() - t0
    updates = float(n) * n * st["matvecs"]  # one node-update = one application of the 9-point operator to one node
    return {"value": updates / (t_tfi + t_smooth), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"single block {n}x{n} (same analytic edges as the GPU workload): TFI + 1 outer iteration gmres/ilu0 rtol 1e-6, "
                      f"{st['matvecs']} operator applications, {t_tfi + t_smooth:.2f} s; the reference is single-threaded",
            "seconds": t_tfi + t_smooth, "tfi_seconds": t_tfi, "krylov_iterations": st["krylov_iterations"]}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return  # the reference is a single-process CPU program; other ranks exit without work
    n = args.ref_size
    vals = []
    for k in range(args.warmup + args.steps):
        s = cpu_baseline_sample(n)
        if k >= args.warmup:
            vals.append(s)
    secs = float(np.mean([v["seconds"] for v in vals]))
    value = float(np.mean([v["value"] for v in vals]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"single_block_{n}x{n} (bounded CPU sample of config 3: synthetic single-block fp64 grid, TFI + elliptic smoothing)",
                       "solver": "gmres+ilu0 rtol 1e-6 (reference defaults), 1 outer iteration per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from turbomesh_b200 import smoothing, synthetic

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
    sweeps = args.sweeps
    solver = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=args.omega, device=local)
    stream = torch.cuda.Stream()                         # the library launches on this stream, so torch events see its kernels
    kind = args.workload or ("single" if world == 1 else "cascade")
    if kind == "single":
        if world != 1:
            raise SystemExit("the single-block workload does not shard; use --workload cascade for N > 1")
        ni = nj = args.size
        spec = synthetic.single_block(ni, nj)
        workload = f"single_block_{ni}x{nj} (config 3: synthetic single-block fp64 grid, TFI + elliptic smoothing)"
        dm = smoothing.DeviceMesh(spec, device=local, stream=stream.cuda_stream, upload=False)
        my_blocks = list(range(len(spec.blocks)))
    else:
        # config 4 in tiling form: one block column (8 blocks of block_ni x block_nj) per GPU; 8 x 8 blocks / 512 Mi nodes at N = 8
        n_bj = args.blocks_per_gpu
        # the passage grows with N (length = N/8) so that the cells stay square: 8 x 8 blocks on a 1 x 0.5 passage at N = 8
        # (the waviness scales with it, so the grid lines have the same inclination at every N)
        spec = synthetic.cascade(world, n_bj, args.block_ni, args.block_nj, length=world / 8.0, ay=0.015 * world / 8.0)
        owner = [bi for bi in range(world) for _ in range(n_bj)]
        workload = (f"cascade_{world}x{n_bj}_blocks_of_{args.block_ni}x{args.block_nj} (config 4: synthetic multi-block cascade passage, "
                    f"{n_bj} blocks per GPU, interface halo exchange once per sweep)")
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

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = smoothing.kernel_launch_count()
    sweep_seconds, sweep_launches = 0.0, 0
    t0 = time.perf_counter()
    stats = None
    ev0.record(stream)
    for _ in range(args.steps):
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
    value = nodes_total * sweeps * args.steps / elapsed

    # ---- time to converged mesh (single block): TFI + FAS multigrid V(3,3) until the mesh changes by <= 1e-10 chord per cycle ----
    ttc = None
    if not args.no_ttc:
        mg = smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, stop_max_update=1e-10, device=local)
        best, cold = None, None
        for _ in range(3 if kind == "single" else 2):
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
        ttc = {"seconds": best[0], "solver_seconds": best[1]["gpu_seconds"], "cycles": best[1]["outer_iterations"],
               "criterion": "max-norm change of the mesh over one V(3,3) cycle <= 1e-10 (chord / passage height are O(1))", "last_max_update": best[1]["last_max_update"],
               "fine_grid_operator_applications": ops, "cold_seconds_incl_hierarchy_setup": cold,
               "solver": "TFI + geometric FAS multigrid over the whole block topology, damped-Jacobi smoother (omega 0.8), Anderson(3) on level-1 samples",
               "equivalent_node_updates_per_s": nodes_total * ops / best[1]["gpu_seconds"],
               "hbm_fraction_of_equivalent_sweeps": nodes_total * ops / best[1]["gpu_seconds"] * BYTES_PER_NODE_UPDATE / 1e9 / (measured_peak()[0] * world),
               "note": "best of %d; includes TFI and begin_smoothing; max over ranks" % (3 if kind == "single" else 2)}

    # ---- end to end through the reference-facing calls with HOST buffers (Block2d.init -> smooth.mesh) ----
    e2e = None
    if kind == "single" and not args.no_e2e:
        e2e = run_e2e(args, spec, solver, torch)
    elif not args.no_e2e:
        e2e = run_e2e_cascade(args, spec, dm, my_blocks, solver, torch, dist, world, nodes_total, barrier)

    peak, peak_src = measured_peak()
    per_launch = sweep_seconds / max(sweep_launches, 1)
    achieved = BYTES_PER_NODE_UPDATE * nodes_local / per_launch / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if (kind == "single" and args.size == 8192) else None, "kernel": "winslow_interior_bulk_kernel<RELAX>",
                "algorithmic_bytes_per_launch": BYTES_PER_NODE_UPDATE * nodes_local, "avg_launch_ms": per_launch * 1e3,
                "peak_source": peak_src + ", sustained copy figure"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "nodes": nodes_total, "sweeps_per_step": sweeps, "omega": args.omega,
                       "step": "TFI of all blocks from device-resident edges + begin_smoothing + sweeps (damped Jacobi, coefficients from the current iterate)",
                       "cache": "inputs (2 x 1.07 GB ping-pong fields per GPU) are larger than the 126 MB L2", "nodes_per_gpu": nodes_local,
                       "halo_exchange": dm.halo_path},
            "roofline": roofline, "clocks": clocks, "gpu_launches": launches, "wall_ms_per_step": wall / args.steps * 1e3,
            "last_max_update": stats["last_max_update"] if stats else None}
    if e2e:
        line["e2e"] = e2e
    line["time_to_converged"] = ttc if ttc else {"seconds": None, "note": "skipped (--no-ttc)"}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1 and kind == "single":
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline_sample(args.ref_size).items() if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    dm.close()
    if world > 1:
        dist.destroy_process_group()


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
    edges = {b: spec.blocks[b].edge_args() for b in my_blocks}
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
    dt = float(t.item())
    return {"value": nodes_total * solver.sweeps_per_iteration / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": dt * 1e3, "api": "tm_mesh_tfi_block (host edges) + tm_mesh_begin_smoothing + tm_mesh_smooth + tm_mesh_download_block (pinned host blocks), per rank",
            "steps": steps, "last_max_update": st["last_max_update"], "note": "bytes per rank"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="turbomesh_b200", choices=["turbomesh_b200", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="single-block edge length (N=1)")
    ap.add_argument("--workload", default=None, choices=["single", "cascade", "cuts"], help="default: single for N=1, cascade for N>1")
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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
