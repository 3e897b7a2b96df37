#!/bin/bash
# scratch job run on the GPU box by tools/_retry.sh (kernel experiments edit this file; tools/micro_scan.py and
# tools/evolve_scan.py are the probes, tools/build_variant.sh builds the variants).  Default: the round-end measurement set.
cd /root/repo
exec bash tools/final_measure.sh
