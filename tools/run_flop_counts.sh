set -x
mkdir -p gpurun_out/flops
python bench.py --write-cpu-sample > gpurun_out/flops/sample.log 2>&1
cp bench_data/cpu_sample_custom_Tc_Qimp.npz gpurun_out/flops/ 
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed.sum
for W in custom_Tc_Qimp fit_lightcurve accretion int16_demo; do
  python bench.py --workload $W --stars 16 --steps 1 --warmup 1 --no-cpu-baseline --no-sharing --no-other-workloads > gpurun_out/flops/bench_$W.json 2> gpurun_out/flops/bench_$W.err || echo FAIL $W
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/flops/flops_$W.csv -k regex:"ds_block_kernel|ds_micro_kernel" -c 8 python bench.py --workload $W --stars 16 --steps 1 --warmup 1 --no-cpu-baseline --no-sharing --no-other-workloads > gpurun_out/flops/ncu_$W.log 2>&1 || echo NCUFAIL $W
done

