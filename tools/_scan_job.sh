#!/bin/bash
cd /root/repo
for v in a b; do
DSTAR_B200_LIB=/root/repo/build/exp/lib_eos_$v.so python tools/micro_scan.py custom_Tc_Qimp 256 eos_variant_$v 2>&1 | grep SCAN
DSTAR_B200_LIB=/root/repo/build/exp/lib_eos_$v.so python tools/micro_scan.py int16_demo 64 eos_variant_$v 2>&1 | grep SCAN
done
