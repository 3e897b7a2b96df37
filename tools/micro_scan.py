"""Coefficient-pass time for one build / launch plan (GPU box):
    [DSTAR_B200_LIB=<so>] [DSTAR_MICRO_SPLIT=0|1] [DSTAR_MICRO_G_EOS=..] [DSTAR_MICRO_G_NU=..] [DSTAR_MICRO_G_FACE=..]
    python tools/micro_scan.py [workload] [stars] [tag]
Prints cells / EOS half / faces in ms (best of 5) and a SHA-256 of the four tables, so that variants can be checked
for bit-identity against each other.  The launch plan is read once per process: one process per variant."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import dstar_b200 as ds
    wl = sys.argv[1] if len(sys.argv) > 1 else "custom_Tc_Qimp"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    tag = sys.argv[3] if len(sys.argv) > 3 else ""
    ds.NScool_init(device=0)
    ctx = ds.Context(0, gaps=("ns", "gc" if wl == "accretion" else "sfb03", "t72"))
    W = bench.prepare_workload(ds, wl, n, 0.05, 0)
    b = ds.Batch(ctx, W["zoff"], W["H"]["ncharged"])
    bench.upload_all(b, W)
    best = [1e30] * 3
    for _ in range(5):
        b.micro_tables(W["mc"], 3)
        for k, w in enumerate((0, 3, 1)):
            best[k] = min(best[k], b.kernel_ms(w))
    h = hashlib.sha256()
    if os.environ.get("DSTAR_SCAN_HASH", "1") != "0":
        t = b.download_tables()
        for k in ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK", "rho_bar1"):
            h.update(np.ascontiguousarray(t[k]).tobytes())
    env = " ".join("%s=%s" % (k[6:], v) for k, v in sorted(os.environ.items()) if k.startswith("DSTAR_") and k != "DSTAR_B200_LIB")
    rc = b.redo_counts()
    print("SCAN %-28s %s stars %d  cells %.3f (eos %.3f nu %.3f)  faces %.3f ms  redo %d+%d of %d  sha %s  [%s]" % (
        tag, wl, W["n_star"], best[0], best[1], best[0] - best[1], best[2], rc[0], rc[1], int(W["zoff"][-1]),
        h.hexdigest()[:12], env), flush=True)
    b.close(); ctx.close(); ds.NScool_shutdown()


if __name__ == "__main__":
    main()
