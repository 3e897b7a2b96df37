#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-other-workloads > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err || tail -5 gpurun_out/r02k_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','curves_per_sec','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['cell_pass'], d['cpu_baseline']['value'])
"
