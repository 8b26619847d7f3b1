timeout 300 python scripts/krylov_probe.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_cuts.py tests/test_gpu_o4h.py -m gpu -q -x 2>&1 | tail -4
