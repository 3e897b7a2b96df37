#!/bin/bash
# build an experimental library (kernel experiments, see tools/micro_scan.py): tools/build_variant.sh <name> [extra nvcc flags...]  -> build/exp/lib_<name>.so
set -e
cd /root/repo/dstar_b200/csrc
name=$1; shift
out=/root/repo/build/exp/$name
mkdir -p $out
FL="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -w"
pids=""
for f in capi.cu atm.cu micro_eos.cu micro_nu.cu micro_face.cu micro_cell.cu host_util.cpp host_nscool.cpp; do
  nvcc $FL "$@" -x cu -c -o $out/$f.o $f & pids="$pids $!"
done
for p in $pids; do wait $p; done
nvcc -shared -o /root/repo/build/exp/lib_$name.so $out/*.o
echo built lib_$name.so
