#!/bin/bash
# experiment job: evolve kernel -- exact branch-free reciprocal vs plain division, one-barrier error norm
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== default: branch-free reciprocal"
timeout 300 python tools/evolve_scan.py custom_Tc_Qimp 148,256,1024 2>&1 | grep "stars\|workload"
timeout 300 python tools/evolve_scan.py accretion 256 2>&1 | grep "stars\|workload"
echo "== plain reciprocal"
DSTAR_B200_LIB=build/exp/lib_plainrcp.so timeout 300 python tools/evolve_scan.py custom_Tc_Qimp 148,256,1024 2>&1 | grep "stars\|workload"
DSTAR_B200_LIB=build/exp/lib_plainrcp.so timeout 300 python tools/evolve_scan.py accretion 256 2>&1 | grep "stars\|workload"
