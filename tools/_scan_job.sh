#!/bin/bash
# experiment job: resident stars per SM for large batches after the evolve changes
cd /root/repo
timeout 600 python tools/evolve_scan.py custom_Tc_Qimp 444,592,1184,2048,4096 2>&1 | grep "stars\|workload"
timeout 900 python tools/evolve_scan.py accretion 444,1024,4096 2>&1 | grep "stars\|workload"
timeout 600 python tools/evolve_scan.py int16_demo 1024 2>&1 | grep "stars\|workload"
