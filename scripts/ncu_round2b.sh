set -x
python scripts/krylov_probe.py t106 > gpurun_out/r2b_probe_t106_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bicgstab_persistent -s 12 -c 1 -f -o gpurun_out/r2b_prof_krylov_persistent python scripts/krylov_probe.py t106 > gpurun_out/r2b_ncu_kp.log 2>&1
python scripts/krylov_probe.py ls89 > gpurun_out/r2b_probe_ls89_plain.log 2>&1
ls -la gpurun_out/*.ncu-rep
CUTS=128 python scripts/krylov_probe.py cuts > gpurun_out/r2b_probe_cuts_plain.log 2>&1 &&
CUTS=128 ncu --set full --clock-control none --import-source on -k regex:krylov_phase_kernel -s 600 -c 6 -f -o gpurun_out/r2b_prof_krylov_phased python scripts/krylov_probe.py cuts > gpurun_out/r2b_ncu_kph.log 2>&1
ls -la gpurun_out/*.ncu-rep
