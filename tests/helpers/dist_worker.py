"""Worker of the multi-process GPU tests (launched with torch.distributed.run, one process per GPU): every rank runs the
distributed solver on its blocks and compares them with a single-process run of the same mesh on its own GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from turbomesh_b200 import smoothing, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args = (2 * world, 2, 65, 33)
    spec = synthetic.cascade(*args)
    mesh0 = synthetic.materialize(spec, smoothing.tfi_block)
    n_bi, n_bj = args[0], args[1]
    owner = [bi * world // n_bi for bi in range(n_bi) for _ in range(n_bj)]
    out = {"rank": rank}
    for name, sol, its in (("relax", smoothing.CudaSolver(method="relax", sweeps_per_iteration=60, omega=0.9, device=local), 2),
                           ("multigrid", smoothing.CudaSolver(method="multigrid", sweeps_per_iteration=3, omega=0.8, device=local), 8),
                           ("picard", smoothing.CudaSolver.tight(device=local), 2)):
        one = mesh0.copy()
        with smoothing.DeviceMesh(one, device=local) as dm:
            dm.begin_smoothing(sol)
            st1 = dm.smooth(its, sol)
            dm.download()
        many = mesh0.copy()
        uid = [smoothing.dist_unique_id() if rank == 0 else None]   # one NCCL id per communicator
        dist.broadcast_object_list(uid, src=0)
        with smoothing.DeviceMesh(many, device=local, owner=owner, rank=rank, n_ranks=world, unique_id=uid[0]) as dm:
            dm.begin_smoothing(sol)
            stn = dm.smooth(its, sol)
            dm.download()
            out["halo_path"] = dm.halo_path
        mine = [b for b in range(len(owner)) if owner[b] == rank]
        out[name] = {"max_diff": max(float(np.abs(one.blocks[b].points - many.blocks[b].points).max()) for b in mine),
                     "update_1": st1["last_max_update"], "update_n": stn["last_max_update"]}
    # The overlapped schedule (rim tiles + exchange on a second stream next to the bulk of the sweep, relax_sweep in
    # csrc/turbomesh_gpu.cu) is only taken when the rim is a small part of the tiles: blocks large enough for that, compared
    # bit for bit with the serial schedule (TM_OVERLAP=0) and, to rounding, with a single process.
    big = (world, 2, 1025, 513)
    spec = synthetic.cascade(*big)
    owner = [bi for bi in range(big[0]) for _ in range(big[1])]
    mine = [b for b in range(len(owner)) if owner[b] == rank]
    sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=25, omega=0.9, device=local)
    res = {}
    for mode in ("1", "0", "single"):
        os.environ["TM_OVERLAP"] = "1" if mode == "single" else mode
        kw = {}
        if mode != "single":
            uid = [smoothing.dist_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            kw = dict(owner=owner, rank=rank, n_ranks=world, unique_id=uid[0])
        with smoothing.DeviceMesh(spec, device=local, upload=False, **kw) as dm:
            for b in (mine if mode != "single" else range(len(owner))):
                dm.tfi_block(b, *spec.blocks[b].edge_args())
            dm.begin_smoothing(sol)
            st = dm.smooth(2, sol)
            res[mode] = ([dm.download_block(b) for b in mine], st["last_max_update"])
    os.environ.pop("TM_OVERLAP", None)
    out["overlap"] = {"bit_identical_to_serial_schedule": all(np.array_equal(a, b) for a, b in zip(res["1"][0], res["0"][0])),
                      "max_diff_to_single_process": max(float(np.abs(a - b).max()) for a, b in zip(res["1"][0], res["single"][0])),
                      "update": res["1"][1], "update_single": res["single"][1]}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("DIST_RESULT " + json.dumps(gathered), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
