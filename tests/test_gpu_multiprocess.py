"""Real multi-process runs (one process per GPU, NCCL bootstrap, halo exchange over NVLink peer memory): needs >= 2 GPUs,
skipped on single-GPU boxes.  The single-GPU tier covers the same rank logic through the in-process emulation
(tests/test_gpu_dist.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_processes_match_a_single_process(gpu_lib, p2p):
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, TM_P2P=p2p)
    port = "29531" if p2p == "1" else "29532"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", port,
           os.path.join(ROOT, "tests", "helpers", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    line = [l for l in res.stdout.splitlines() if l.startswith("DIST_RESULT ")]
    assert res.returncode == 0 and line, res.stdout[-2000:] + res.stderr[-2000:]
    for r in json.loads(line[0][len("DIST_RESULT "):]):
        assert r["halo_path"].startswith("nvlink peer memory" if p2p == "1" else "nccl"), r
        assert r["relax"]["max_diff"] <= 1e-14 and r["multigrid"]["max_diff"] <= 1e-13 and r["picard"]["max_diff"] <= 1e-10, r
        # the overlapped rim / bulk schedule on blocks large enough to take it (1025 x 513)
        assert r["overlap"]["bit_identical_to_serial_schedule"] and r["overlap"]["max_diff_to_single_process"] <= 1e-14, r
        assert r["overlap"]["update"] > 0
