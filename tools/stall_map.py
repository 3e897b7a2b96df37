"""Where a kernel's warp-stall samples fall in the source: joins the SASS page of an ncu report
    ncu -i rep.ncu-rep --page source --csv --print-source sass --launch-count 1 [--launch-skip N] > k.csv
with the line table of the same cubin (nvdisasm -gi -c kernel.cubin, restricted to the kernel) by instruction order, and
buckets the samples by the call site inside an anchor function (file, first line, last line):
    python tools/stall_map.py k.csv k.sass eos.cuh 846 960"""
import collections
import csv
import os
import re
import sys


def sass_chains(path):
    chains, chain, fresh, ops = [], [], True, []
    for l in open(path):
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), m.group(1)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            fresh = True
            chains.append(list(chain))
            ops.append(m.group(2))
    return chains, ops


def main(csv_path, sass_path, afile, lo, hi):
    lo, hi = int(lo), int(hi)
    rows = list(csv.reader(open(csv_path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ci = {k: i for i, k in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) >= len(hdr)]
    chains, ops = sass_chains(sass_path)
    if len(chains) != len(data):
        print("warning: %d instructions in the report, %d in the disassembly" % (len(data), len(chains)))
    keys = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_math", "stall_not_selected", "stall_lg", "stall_short_sb",
            "stall_no_inst", "stall_selected", "stall_mio", "stall_branch_resolving"]
    agg = collections.defaultdict(lambda: collections.Counter())
    src = {}
    for r, ch in zip(data, chains):
        key = "(outside)"
        for f, ln, full in ch:
            if f == afile and lo <= ln <= hi:
                key = ln
                if ln not in src:
                    try:
                        src[ln] = open(full).read().split("\n")[ln - 1].strip()[:70]
                    except OSError:
                        src[ln] = ""
        a = agg[key]
        a["samples"] += int(r[ci["# Samples"]] or 0)
        a["inst"] += int(r[ci["Instructions Executed"]] or 0)
        for k in keys:
            a[k] += int(r[ci[k]] or 0)
    tot = sum(a["samples"] for a in agg.values())
    print("total samples %d" % tot)
    print("%9s %6s %9s " % ("line", "smpl%", "warpinst") + " ".join("%7s" % k[6:13] for k in keys) + "  source")
    for k in sorted(agg, key=lambda k: (isinstance(k, str), k)):
        a = agg[k]
        if a["samples"] < tot * 0.002:
            continue
        print("%9s %6.1f %9d " % (k, 100.0 * a["samples"] / tot, a["inst"]) + " ".join("%7.1f" % (100.0 * a[q] / tot) for q in keys) +
              "  " + src.get(k, ""))


if __name__ == "__main__":
    main(*sys.argv[1:6])
