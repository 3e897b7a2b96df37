#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/evolve_scan.py custom_Tc_Qimp 148,256,1024 2>&1 | grep -E "stars"
python tools/evolve_scan.py accretion 256 2>&1 | grep -E "stars"
