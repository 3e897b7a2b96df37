"""Evolve-kernel time versus batch size and CTAs per SM (GPU box): python tools/evolve_scan.py [workload]."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import dstar_b200 as ds
    wl = sys.argv[1] if len(sys.argv) > 1 else "custom_Tc_Qimp"
    sizes = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "64,128,148,256,296,444,592,1024,2048,4096".split(","))]
    ds.NScool_init(device=0)
    ctx = ds.Context(0, gaps=("ns", "gc" if wl == "accretion" else "sfb03", "t72"))
    base = bench.prepare_workload(ds, wl, 64 if wl != "custom_Tc_Qimp" else 256, 0.05, 0)
    # tables once, on the distinct stars; larger batches reuse them through table sharing
    print("workload", wl, "distinct", base["n_distinct"], "nz", int(np.median(base["nz"])))
    for n in sizes:
        reps = max(1, -(-n // base["n_distinct"]))
        nz = base["nz"][:min(n, base["n_distinct"])]
        sel = dict(base)
        if n < base["n_distinct"]:
            picks = list(range(n))
            S = bench.sample_from_inputs(base["H"], base["zoff"], picks)
            H = {k: S[k] for k in S if k != "zoff"}
            sel.update(H=H, nz=np.diff(S["zoff"]), zoff=S["zoff"], heat_arr=base["heat_arr"][:n], n_star=n, T_atm_arr=base["T_atm_arr"][:n])
        else:
            nz2, zoff2, H2, heat2 = bench.tile_inputs(base["nz"], base["H"], base["heat_arr"], reps)
            sel.update(H=H2, nz=nz2, zoff=zoff2, heat_arr=heat2, n_star=len(nz2), T_atm_arr=np.full(len(nz2), base["T_atm"]))
        b = ds.Batch(ctx, sel["zoff"], sel["H"]["ncharged"])
        bench.upload_all(b, sel)
        cell, face = bench.sharing_owners(sel)
        b.set_table_sharing(1, cell); b.set_table_sharing(2, face)
        b.micro_tables(sel["mc"], 3)
        row = []
        for occ in (1, 2, 3, 4, 0):
            b.set_evolve_occupancy(occ)
            ts = []
            for _ in range(3):
                b.reset_state(); b.evolve(sel["ec"], 0); ts.append(b.kernel_ms(2))
            row.append(min(ts))
        res = b.download_results(full=False)
        print("stars %5d  evolve ms at 1/2/3/4 CTAs per SM, auto: %s  steps/star %.1f converged %s" % (
            sel["n_star"], " ".join("%8.3f" % t for t in row), float(np.mean(res["counters"][:, 2])), bool(np.all(res["idid"] >= 0))), flush=True)
        b.close()
    ctx.close(); ds.NScool_shutdown()


if __name__ == "__main__":
    main()
