#!/bin/bash
# Round-end single-GPU measurement set (GPU box): tests, executed-flop counts, launch list, ncu captures, bench lines.
# gpurun copies back at most 64 MiB: the reports are exported to CSV on the box and only the EOS one is kept.
cd /root/repo
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/run_flop_counts.sh > gpurun_out/flops_run.log 2>&1
O="--steps 2 --warmup 1 --no-cpu-baseline --no-sharing --no-other-workloads"
python bench.py $O > gpurun_out/r02_pre.json 2> gpurun_out/r02_pre.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $O > gpurun_out/r02_launches_ncu.log 2>&1 || echo LAUNCHLIST_FAIL
O1="--steps 1 --warmup 0 --no-cpu-baseline --no-sharing --no-other-workloads"
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
exp() {  # exp <name> [keep|nosource]: raw + source pages as CSV, report removed unless kept
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  if [ "$2" != "nosource" ]; then ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass --launch-count 1 > gpurun_out/$1_source.csv 2>/dev/null; fi
  if [ "$2" != "keep" ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -f -k regex:"ds_block_kernel|ds_evolve_kernel|ds_zone_aux|ds_eos_|ds_interp1" -c 14 -o gpurun_out/r02_full_default python bench.py $O1 > gpurun_out/r02_full_default.log 2>&1 || echo FULL_DEFAULT_FAIL
exp r02_full_default nosource
timeout 900 $N -k regex:"ds_block_kernel<\(int\)1, \(bool\)1" -c 1 -o gpurun_out/r02_full_eos python bench.py $O1 > gpurun_out/r02_full_eos.log 2>&1 || echo FULL_EOS_FAIL
exp r02_full_eos keep
timeout 900 $N -k regex:ds_evolve_kernel -c 1 -o gpurun_out/r02_full_evolve python bench.py $O1 > gpurun_out/r02_full_evolve.log 2>&1 || echo FULL_EVOLVE_FAIL
exp r02_full_evolve
timeout 900 $N -k regex:"ds_block_kernel<\(int\)3" -c 1 -o gpurun_out/r02_full_accretion_face python bench.py --workload accretion --stars 64 $O1 > gpurun_out/r02_full_accretion.log 2>&1 || echo FULL_ACC_FAIL
exp r02_full_accretion_face
timeout 900 $N -k regex:"ds_block_kernel<\(int\)1, \(bool\)1" -c 1 -o gpurun_out/r02_full_int16_eos python bench.py --workload int16_demo --stars 64 $O1 > gpurun_out/r02_full_int16.log 2>&1 || echo FULL_INT16_FAIL
exp r02_full_int16_eos
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err
ls -la gpurun_out/; du -sh gpurun_out
