#!/bin/bash
# experiment job: interpolation CTA size (zones per CTA 1 / 2 / 4), EOS warps per CTA after the leaner divisions
cd /root/repo
timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 default_g4_w24 2>&1 | grep SCAN
for v in ig2 ig1 w16 w20 w28; do DSTAR_B200_LIB=build/exp/lib_$v.so timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 $v 2>&1 | grep SCAN; done
for v in w16 w20; do DSTAR_B200_LIB=build/exp/lib_$v.so timeout 300 python tools/micro_scan.py int16_demo 64 $v 2>&1 | grep SCAN; done
timeout 300 python tools/micro_scan.py int16_demo 64 default 2>&1 | grep SCAN
