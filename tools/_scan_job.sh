#!/bin/bash
# experiment job: evolve kernel -- bounds-cached table segments, zone 0 on the common path
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/evolve_scan.py custom_Tc_Qimp 148,256,1024 2>&1 | grep "stars\|workload"
timeout 300 python tools/evolve_scan.py accretion 256 2>&1 | grep "stars\|workload"
