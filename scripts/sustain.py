"""Burst vs sustained sweep time (power cap / clocks), single block."""
import os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbomesh_b200 import smoothing, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
spec = synthetic.single_block(n, n)
dm = smoothing.DeviceMesh(spec, upload=False)
dm.tfi_block(0, *spec.blocks[0].edge_args())
def clk():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for sweeps in (20, 100, 500, 2000, 20):
    sol = smoothing.CudaSolver(method="relax", sweeps_per_iteration=sweeps, omega=0.9)
    dm.begin_smoothing(sol)
    st = dm.smooth(1, sol)
    per = st["gpu_seconds"] / sweeps
    print(f"sweeps {sweeps:5d}: {per*1e3:.4f} ms/sweep  {n*n*32/per/1e9:7.1f} GB/s   [{clk()}]")
    time.sleep(0.5)
