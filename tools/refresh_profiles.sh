#!/bin/bash
# After tools/final_measure.sh has run on the GPU box and gpurun_out/ has been merged back: copy the bench lines and the
# launch list into profiles/, recount the executed flops, and regenerate the stall maps of the EOS block kernel and the
# evolve kernel (SASS with line tables from the objects under build/obj, which must be the build that ran).
# Prints the ncu tables (tools/ncu_summary.py) for pasting into profiles/r02_ncu_full_summary.md.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out
cp $G/r02_launches.csv profiles/r02_launches.csv
cp $G/r02_bench_final.json profiles/r02_bench_256stars.json
cp $G/r02_bench_reference_final.json profiles/r02_bench_reference_arm.json
F=$G/flops
python tools/count_flops.py profiles/r02_flop_counts.json custom_Tc_Qimp=$F/flops_custom_Tc_Qimp.csv:$F/bench_custom_Tc_Qimp.json \
  fit_lightcurve=$F/flops_fit_lightcurve.csv:$F/bench_fit_lightcurve.json accretion=$F/flops_accretion.csv:$F/bench_accretion.json \
  int16_demo=$F/flops_int16_demo.csv:$F/bench_int16_demo.json > /dev/null
T=$(mktemp -d)
( cd $T && cuobjdump -xelf all "$OLDPWD/build/obj/micro_eos.cu.o" > /dev/null && cuobjdump -xelf all "$OLDPWD/build/obj/capi.cu.o" > /dev/null \
  && nvdisasm -gi -c micro_eos.sm_100a.cubin > eos_all.sass && nvdisasm -gi -c capi.sm_100a.cubin > capi.sass )
python - "$T" <<'PY'
import sys
T = sys.argv[1]
def cut(path, name, out):
    L = open(path).read().split('\n')
    idx = [i for i, l in enumerate(L) if l.startswith('\t.section\t.text.')]
    for n, i in enumerate(idx):
        if name in L[i]:
            j = idx[n + 1] if n + 1 < len(idx) else len(L)
            open(out, 'w').write('\n'.join(L[i:j]))
            return
    raise SystemExit('kernel not found: ' + name)
cut(T + '/eos_all.sass', '_Z15ds_block_kernelILi1ELb1EE', T + '/eos_blk.sass')
cut(T + '/capi.sass', '_Z16ds_evolve_kernelILi256ELi1EE', T + '/evolve.sass')
PY
E0=$(grep -n "^DS_D void ds_eval_crust_eos" dstar_b200/csrc/eos.cuh | cut -d: -f1)
E1=$(awk -v s=$E0 'NR>s && /^}/ {print NR; exit}' dstar_b200/csrc/eos.cuh)
K0=$(grep -n "ds_evolve_kernel(DsEvolveArgs" dstar_b200/csrc/kernels_evolve.cuh | cut -d: -f1)
K1=$(awk -v s=$K0 'NR>s && /^}/ {print NR; exit}' dstar_b200/csrc/kernels_evolve.cuh)
python tools/stall_map.py $G/r02_full_eos_source.csv $T/eos_blk.sass eos.cuh $E0 $E1 2>&1 | awk 'NR<=2 || $2+0>=0.25' | cut -c1-205 > $G/eos_map.txt
python tools/stall_map.py $G/r02_full_evolve_source.csv $T/evolve.sass kernels_evolve.cuh $K0 $K1 2>&1 | awk 'NR<=2 || $2+0>=1.0' | cut -c1-205 > $G/ev_map.txt
for f in default evolve accretion_face int16_eos; do echo "== $f"; python tools/ncu_summary.py $G/r02_full_${f}_raw.csv; done
echo "== stall maps: $G/eos_map.txt $G/ev_map.txt"
rm -rf $T
