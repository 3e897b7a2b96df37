#!/bin/bash
# _retry.sh <logname> <timeout> : run tools/_scan_job.sh on the GPU box, retrying while the pod is busy
for i in 1 2 3 4 5 6 7 8; do
  gpurun --timeout $2 -- "bash tools/_scan_job.sh > gpurun_out/$1.log 2>&1; cat gpurun_out/$1.log" > gpurun_out/$1_call.log 2>&1
  if grep -q "status=ok\|status=fail\|status=timeout" gpurun_out/$1_call.log; then break; fi
  sleep 90
done
