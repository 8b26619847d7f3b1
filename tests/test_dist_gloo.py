"""N > 1 host logic on the CPU: two gloo ranks each build their partition plan (tm_dist_plan, no GPU) and check that
what one rank sends is exactly what the other expects, then route a field through gloo the way the GPU ranks route it
through NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist

        from turbomesh_b200 import build, smoothing, synthetic

        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        build.build()
        spec = synthetic.cascade(4, 2, 21, 13)
        nb = len(spec.blocks)
        owner = [b * world // nb for b in range(nb)]
        plan = smoothing.dist_plan(spec, owner, rank, world)
        plans = [None] * world
        dist.all_gather_object(plans, {k: (v if not isinstance(v, list) else [a.tolist() for a in v]) for k, v in plan.items()})
        # what I send to p is what p expects from me, in the same order
        for p in range(world):
            if p != rank:
                assert plans[p]["ghost_ids"][rank] == plan["send_ids"][p].tolist()
                assert plan["ghost_ids"][p].tolist() == plans[p]["send_ids"][rank]
        assert sum(pl["n_own"] for pl in plans) == sum(b.size[0] * b.size[1] for b in spec.blocks)
        assert plan["n_ghost"] > 0 and len(plan["ghost_ids"][rank]) == 0
        # route a field whose value is the global node id: afterwards every ghost slot must hold its own id
        offs = np.cumsum([0] + [b.size[0] * b.size[1] for b in spec.blocks])
        field = {int(g): float(g) for b in range(nb) if owner[b] == rank for g in range(offs[b], offs[b + 1])}
        reqs, recv = [], {}
        for p in range(world):
            if p == rank:
                continue
            if len(plan["send_ids"][p]):
                buf = torch.tensor([field[int(g)] for g in plan["send_ids"][p]], dtype=torch.float64)
                reqs.append(dist.isend(buf, p))
            if len(plan["ghost_ids"][p]):
                recv[p] = torch.empty(len(plan["ghost_ids"][p]), dtype=torch.float64)
                reqs.append(dist.irecv(recv[p], p))
        for r in reqs:
            r.wait()
        for p, buf in recv.items():
            assert np.array_equal(buf.numpy(), plan["ghost_ids"][p].astype(np.float64))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))


def test_partition_plans_are_consistent_across_gloo_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results


def test_partition_plan_single_rank_has_no_ghosts(gpu_lib):
    sys.path.insert(0, ROOT)
    from turbomesh_b200 import smoothing, synthetic

    spec = synthetic.cascade(2, 2, 12, 9)
    plan = smoothing.dist_plan(spec, [0, 0, 0, 0], 0, 1)
    assert plan["n_ghost"] == 0 and plan["n_synth"] == 0 and plan["n_send"] == 0
    assert plan["n_own"] == 4 * 12 * 9
    plan2 = smoothing.dist_plan(spec, [0, 0, 1, 1], 0, 2)
    assert plan2["n_own"] == 2 * 12 * 9 and plan2["n_ghost"] > 0
