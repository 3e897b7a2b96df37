#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/evolve_scan.py custom_Tc_Qimp 256,444,592 2>&1 | grep -E "stars"
python tools/micro_scan.py custom_Tc_Qimp 256 default 2>&1 | grep SCAN
DSTAR_STACK_LIMIT=0 python tools/micro_scan.py custom_Tc_Qimp 256 stack0 2>&1 | grep SCAN
python bench.py --steps 10 --warmup 3 --no-other-workloads > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err || tail -5 gpurun_out/r02i_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','curves_per_sec','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['cell_pass'])
"
