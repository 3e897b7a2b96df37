#!/bin/bash
# experiment job: persistent interpolation kernels (CTAs per SM 4 / 3 / 2) against the one-round form
cd /root/repo
timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 ic4 2>&1 | grep SCAN
for v in ic3 ic2 ione; do DSTAR_B200_LIB=build/exp/lib_$v.so timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 $v 2>&1 | grep SCAN; done
timeout 300 python tools/micro_scan.py accretion 256 ic4 2>&1 | grep SCAN
timeout 300 python tools/micro_scan.py int16_demo 64 ic4 2>&1 | grep SCAN
timeout 300 python tools/micro_scan.py fit_lightcurve 256 ic4 2>&1 | grep SCAN
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
