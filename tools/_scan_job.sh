#!/bin/bash
# round-2 final single-GPU measurement set (block-scheduled kernels): tests, executed-flop counts, launch list, ncu captures
cd /root/repo
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/run_flop_counts.sh > gpurun_out/flops_run.log 2>&1
O="--steps 2 --warmup 1 --no-cpu-baseline --no-sharing --no-other-workloads"
python bench.py $O > gpurun_out/r02_pre.json 2> gpurun_out/r02_pre.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $O > gpurun_out/r02_launches_ncu.log 2>&1 || echo LAUNCHLIST_FAIL
O1="--steps 1 --warmup 0 --no-cpu-baseline --no-sharing --no-other-workloads"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ds_block_kernel|ds_evolve_kernel|ds_zone_aux|ds_eos_|ds_interp1" -c 14 -f -o gpurun_out/r02_full_default python bench.py $O1 > gpurun_out/r02_full_default.log 2>&1 || echo FULL_DEFAULT_FAIL
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ds_block_kernel<\(int\)3" -c 1 -f -o gpurun_out/r02_full_accretion_face python bench.py --workload accretion --stars 64 $O1 > gpurun_out/r02_full_accretion.log 2>&1 || echo FULL_ACC_FAIL
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ds_block_kernel<\(int\)1, \(bool\)1" -c 1 -f -o gpurun_out/r02_full_int16_eos python bench.py --workload int16_demo --stars 64 $O1 > gpurun_out/r02_full_int16.log 2>&1 || echo FULL_INT16_FAIL
ls -la gpurun_out/*.ncu-rep
