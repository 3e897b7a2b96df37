#!/usr/bin/env python3
"""Write superfluid critical-temperature tables in the reference's Tc_data file format.

The reference's superfluid/data/Tc_data.txz is a git-LFS stub, so the gap tables are rebuilt:
  * 'gc' (n 1S0), 'ccdk93' (p 1S0), 'ao85' (n 3P2): resampled from the reference's own golden
    output superfluid/test/sample_output.data (40 samples of Tc(kF), 6 decimals in GK) --
    committed copy: tests/golden/reference/superfluid.sample_output.data.
  * every other name in NScool/defaults/controls.defaults: the phenomenological form of
    Ho, Elshamouty, Heinke & Potekhin (2015, PRC 91, 015806), eq. (2),
        Delta(kF) = D0 (kF-k0)^2/((kF-k0)^2+k1) * (kF-k2)^2/((kF-k2)^2+k3),  k0 < kF < k2,
    Tc = 0.5669 Delta/kB (singlet), Tc = 0.8418 Delta/kB (triplet), with their Table II
    parameters where available and nearby stand-ins otherwise (SYNTHETIC, flagged below).
File format (superfluid/private/load_sf.f:40-58): 3 header lines, 'nv kF_min kF_max',
then nv kF values, then nv Tc values [K].
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "dstar_b200", "data", "Tc_data")
GOLD = os.path.join(ROOT, "tests", "golden", "reference", "superfluid.sample_output.data")
MEV_K = 1.160451812e10

# name: (D0 [MeV], k0, k1, k2, k3, exact_from_Ho2015)
P1 = {
    "ao85": (14.0, 0.15, 0.22, 1.05, 3.8, True),
    "bcll92": (120.0, 0.0, 9.0, 1.3, 1.8, True),
    "ccy72-ms": (76.0, 0.0, 12.0, 1.1, 0.8, False),
    "ccy72-ps": (150.0, 0.0, 25.0, 1.2, 1.5, False),
    "eeho96": (4.5, 0.0, 0.57, 1.2, 0.35, True),
    "ns": (17.0, 0.0, 2.9, 0.8, 0.08, False),
    "t73": (48.0, 0.15, 2.1, 1.2, 2.8, True),
}
N1 = {
    "ccdk93": (127.0, 0.18, 4.5, 1.08, 1.1, True),
    "gipsf08": (8.8, 0.18, 0.1, 1.2, 0.6, True),
    "gipsf08-2": (8.8, 0.18, 0.1, 1.3, 0.6, False),
    "gipsf08-3": (8.8, 0.18, 0.1, 1.5, 0.6, False),
    "sfb03": (45.0, 0.1, 4.5, 1.55, 2.5, True),
    "wap": (69.0, 0.15, 3.0, 1.4, 3.0, True),
}
N3 = {
    "hgrr70": (4.8, 1.07, 1.8, 3.2, 2.0, False),
    "t72": (4.8, 1.07, 1.8, 3.2, 2.0, False),
    "bcll92": (0.068, 1.28, 0.1, 2.37, 0.02, False),
}


def ho_tc(par, triplet, n=64):
    D0, k0, k1, k2, k3, _ = par
    kF = np.linspace(k0, k2, n)
    d = D0 * (kF - k0) ** 2 / ((kF - k0) ** 2 + k1) * (kF - k2) ** 2 / ((kF - k2) ** 2 + k3)
    fac = 0.8418 if triplet else 0.5669
    return kF, fac * d * MEV_K


def write(name, code, kF, Tc, note):
    path = os.path.join(OUT, "%s_%s.data" % (name, code))
    with open(path, "w") as f:
        f.write("# %s %s critical temperature table (dstar-b200 rebuild)\n" % (name, code))
        f.write("# %s\n" % note)
        f.write("# nv kF_min kF_max / kF [fm^-1] / Tc [K]\n")
        f.write("%d %.10e %.10e\n" % (len(kF), kF[0], kF[-1]))
        f.write(" ".join("%.10e" % v for v in kF) + "\n")
        f.write(" ".join("%.10e" % v for v in Tc) + "\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    rows = [l.split() for l in open(GOLD).read().splitlines()[2:] if l.strip()]
    g = np.array(rows, dtype=float)
    k = g[:, 0]
    # the golden prints k with 3 decimals; the test program uses k = 3.6*(i-1)/39
    k = 3.6 * np.arange(40) / 39.0

    def trimmed(col):
        tc = g[:, col] * 1.0e9
        nz = np.nonzero(tc > 0)[0]
        lo, hi = max(nz[0] - 1, 0), min(nz[-1] + 1, len(k) - 1)
        return k[lo:hi + 1], tc[lo:hi + 1]

    note = "resampled from the reference golden superfluid/test/sample_output.data"
    kk, tc = trimmed(1); write("ccdk93", "p1", kk, tc, note)
    kk, tc = trimmed(2); write("gc", "n1", kk, tc, note)
    kk, tc = trimmed(3); write("ao85", "n3", kk, tc, note)
    for code, tab, trip in (("p1", P1, False), ("n1", N1, False), ("n3", N3, True)):
        for name, par in tab.items():
            if os.path.exists(os.path.join(OUT, "%s_%s.data" % (name, code))) and name in ("ccdk93", "gc", "ao85") \
                    and (name, code) in (("ccdk93", "p1"), ("gc", "n1"), ("ao85", "n3")):
                continue
            kF, Tc = ho_tc(par, trip)
            src = "Ho et al. (2015) eq.(2), Table II" if par[5] else "SYNTHETIC stand-in, Ho et al. (2015) functional form"
            write(name, code, kF, Tc, src)
    print("wrote gap tables to", OUT)


if __name__ == "__main__":
    sys.exit(main())
