#!/bin/bash
cd /root/repo
for B in 3 7; do DSTAR_MICRO_BLOCKS=$B timeout 300 python tools/micro_scan.py accretion 256 blocks$B 2>&1 | grep SCAN; done
DSTAR_MICRO_BLOCKS=7 timeout 300 python tools/micro_scan.py custom_Tc_Qimp 256 blocks7 2>&1 | grep SCAN
