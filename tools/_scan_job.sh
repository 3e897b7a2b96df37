#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r02_bench_256stars.json 2> gpurun_out/r02_bench_256stars.err || tail -5 gpurun_out/r02_bench_256stars.err
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err || tail -5 gpurun_out/r02_bench_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke()"
tail -c 400 gpurun_out/r02_bench_reference_arm.json
