"""FP64 operations EXECUTED per microphysics evaluation, from hardware counters (the mechanical count).

Input: for each workload one ncu metrics pass over a short bench run,
    ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed.sum --clock-control none --csv \
        --log-file flops_<workload>.csv -k regex:ds_micro_kernel python bench.py --workload <workload> --stars 16 ...
and the JSON line that run printed (for the zone count).  flop = dadd + dmul + 2 dfma per thread, divided by the
evaluations of the launch (zones x 128 knots).  Divisions, square roots and the elementary functions are counted as the
FP64 operations they expand to (that is what executes; SURVEY's 6.0 k / 1.6 k estimate counted them by weight).

Usage: python tools/count_flops.py OUT.json workload=flops.csv:bench.json [...]
"""
import csv
import json
import sys


def read(csv_path):
    rows = [r for r in csv.reader(open(csv_path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    k_i, m_i, v_i = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
    id_i = H.index("ID")
    launches = {}
    for r in rows[hdr + 1:]:
        if len(r) <= v_i:
            continue
        d = launches.setdefault(int(r[id_i]), {"kernel": r[k_i]})
        d[r[m_i]] = float(r[v_i].replace(",", ""))
    return [launches[k] for k in sorted(launches)]


def main():
    out_path = sys.argv[1]
    out = {}
    for spec in sys.argv[2:]:
        wl, rest = spec.split("=")
        csv_path, json_path = rest.split(":")
        line = json.loads([l for l in open(json_path) if l.startswith("{")][-1])
        zones = line["config"]["zones_total_per_gpu"]
        evals = zones * 128.0
        ent = {"source": "ncu hardware counters, %s, %d zones x 128 knots per launch (tools/count_flops.py)" % (csv_path.split("/")[-1], zones)}
        # ds_block_kernel<PART, FAST> (or ds_micro_kernel<PART, G, FAST> with DSTAR_MICRO_BLOCKS=0): PART 1 = EOS half,
        # 2 = neutrino half (cells), 3 = faces (kernels_micro.cuh); the interpolation kernels (~30 flop per evaluation) are
        # not counted
        launches = read(csv_path)
        for kind, part in (("eos", 1), ("nu", 2), ("face", 3)):
            pats = ("ds_block_kernel<%d," % part, "ds_micro_kernel<%d," % part)
            ls = [l for l in launches if any(p_ in l["kernel"].replace(" ", "").replace("(bool)", "").replace("(int)", "") for p_ in pats)]
            ls = [l for l in ls if l.get("smsp__thread_inst_executed.sum", 0.0) > 1.0e6]   # skip the (empty) plain-division passes
            if not ls:
                continue
            l = ls[0]
            dadd = l["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
            dmul = l["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
            dfma = l["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
            ent[kind] = {"flop_per_eval": (dadd + dmul + 2.0 * dfma) / evals, "dadd": dadd / evals, "dmul": dmul / evals,
                         "dfma": dfma / evals, "inst_per_eval": l.get("smsp__thread_inst_executed.sum", 0.0) / evals,
                         "launches_seen": len(ls)}
        if "eos" in ent and "nu" in ent:
            ent["cell"] = {k: ent["eos"][k] + ent["nu"][k] for k in ("flop_per_eval", "dadd", "dmul", "dfma", "inst_per_eval")}
        out[wl] = ent
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
