# ncu evidence of round 2 (one gpurun call; every command has run plainly first)
set -x
M="gpu__time_duration.sum"
python scripts/probe_passages.py 48 10 > gpurun_out/r2_probe_passages_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:winslow_interior_bulk -s 12 -c 3 -f -o gpurun_out/r2_prof_sweep_passages python scripts/probe_passages.py 48 10 > gpurun_out/r2_ncu_sweep.log 2>&1
python scripts/probe.py 8192 5 > gpurun_out/r2_probe_single_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tfi_kernel -s 2 -c 2 -f -o gpurun_out/r2_prof_tfi python scripts/probe.py 8192 5 > gpurun_out/r2_ncu_tfi.log 2>&1
python scripts/krylov_probe.py t106 > gpurun_out/r2_probe_t106_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bicgstab_persistent -s 12 -c 1 -f -o gpurun_out/r2_prof_krylov_persistent python scripts/krylov_probe.py t106 > gpurun_out/r2_ncu_kp.log 2>&1
CUTS=128 python scripts/krylov_probe.py cuts > gpurun_out/r2_probe_cuts_plain.log 2>&1 &&
CUTS=128 ncu --set full --clock-control none --import-source on -k regex:krylov_phase -s 300 -c 6 -f -o gpurun_out/r2_prof_krylov_phased python scripts/krylov_probe.py cuts > gpurun_out/r2_ncu_kph.log 2>&1
python bench.py --steps 2 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_plain_for_ncu.log 2>&1 &&
ncu --metrics $M --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
