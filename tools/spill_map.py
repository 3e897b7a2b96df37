"""Static map of register-spill traffic of a kernel: counts LDL/STL (local loads/stores) in its SASS per call site
of an anchor function.
usage: nvdisasm -gi -c kernel.cubin > k.sass
       python tools/spill_map.py k.sass eos.cuh 844 941      # bucket by the line of ds_eval_crust_eos's body
The coefficient kernels are straight-line code executed once per lane, so static counts ~ dynamic counts
(except inside the species loop of ion_mixture)."""
import collections
import os
import re
import sys


def main(p, afile, lo, hi):
    lo, hi = int(lo), int(hi)
    chain = []
    fresh = True
    ninst, ldl, stl = collections.Counter(), collections.Counter(), collections.Counter()
    src = {}
    for l in open(p):
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), m.group(1)))
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", l):
            fresh = True
            key = "(outside)"
            for f, ln, full in chain:
                if f == afile and lo <= ln <= hi:
                    key = ln
                    if ln not in src:
                        try:
                            src[ln] = open(full).read().split("\n")[ln - 1].strip()[:90]
                        except OSError:
                            src[ln] = ""
            ninst[key] += 1
            if re.search(r"\bLDL", l): ldl[key] += 1
            if re.search(r"\bSTL", l): stl[key] += 1
    print(f"{'line':>9s} {'inst':>7s} {'LDL':>5s} {'STL':>5s}  source")
    for k in sorted(ninst, key=lambda k: (isinstance(k, str), k)):
        if ninst[k] < 30 and ldl[k] + stl[k] == 0: continue
        print(f"{str(k):>9s} {ninst[k]:7d} {ldl[k]:5d} {stl[k]:5d}  {src.get(k, '')}")
    print(f"{'total':>9s} {sum(ninst.values()):7d} {sum(ldl.values()):5d} {sum(stl.values()):5d}")


if __name__ == "__main__":
    main(*sys.argv[1:5])
