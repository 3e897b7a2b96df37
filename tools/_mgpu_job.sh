#!/bin/bash
# round-2 multi-GPU lines (8-GPU box): weak scaling 2/4/8, strong scaling 2048 stars over 1/2/4/8, config[4] at 1024 stars per GPU,
# the stack-limit question under NCCL
cd /root/repo
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29600
for N in 2 4 8; do
  P=$((P+1)); $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err || echo FAIL weak $N
done
P=$((P+1)); DSTAR_STACK_LIMIT=0 $TR --nproc-per-node 2 --master-port $P bench.py --gpus 2 --steps 20 --warmup 3 --no-sharing > gpurun_out/r02_bench_2gpu_stack0.json 2> gpurun_out/r02_bench_2gpu_stack0.err || echo FAIL stack0
python bench.py --total-stars 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads --no-sharing > gpurun_out/r02_strong_1gpu.json 2> gpurun_out/r02_strong_1gpu.err || echo FAIL strong 1
for N in 2 4 8; do
  P=$((P+1)); $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --total-stars 2048 --steps 10 --warmup 3 --no-sharing > gpurun_out/r02_strong_${N}gpu.json 2> gpurun_out/r02_strong_${N}gpu.err || echo FAIL strong $N
done
P=$((P+1)); $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --workload int16_demo --stars 1024 --steps 3 --warmup 3 --no-sharing > gpurun_out/r02_int16_8gpu.json 2> gpurun_out/r02_int16_8gpu.err || echo FAIL int16
for f in gpurun_out/r02_bench_*gpu*.json gpurun_out/r02_strong_*.json gpurun_out/r02_int16_8gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[1], d['n_gpus'], d['scaling'], 'ms', round(d['ms_per_step'],3), 'value %.4g'%d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), d['roofline']['kernel_ms'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
