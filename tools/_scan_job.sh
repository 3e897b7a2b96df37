#!/bin/bash
cd /root/repo
timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 blocks_w28 2>&1 | grep SCAN
for W in 24 32; do DSTAR_B200_LIB=/root/repo/build/exp/lib_w$W.so timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 blocks_w$W 2>&1 | grep SCAN; done
timeout 300 python tools/micro_scan.py accretion 256 blocks_w28 2>&1 | grep SCAN
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
