"""Shared helpers for the test-suite (file readers for fixtures; no product or oracle code)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference")
GAPS = os.path.join(ROOT, "dstar_b200", "data", "Tc_data")


def read_rows(path, skip=0, after=None, ncol=None):
    lines = open(path).read().split("\n")
    if after is not None:
        for i, l in enumerate(lines):
            if after in l:
                lines = lines[i + 1:]
                break
    rows = []
    for l in lines[skip:]:
        p = l.replace("-1", "E-1") .split() if False else l.split()
        if ncol is not None and len(p) != ncol:
            continue
        try:
            rows.append([_flt(v) for v in p])
        except ValueError:
            continue
    return rows


def _flt(v):
    # Fortran drops the 'E' for 3-digit exponents: 1.496843-107
    try:
        return float(v)
    except ValueError:
        for i in range(len(v) - 1, 0, -1):
            if v[i] in "+-" and v[i - 1].isdigit():
                return float(v[:i] + "E" + v[i:])
        raise


def load_gap_file(name, code):
    """Tc_data format (superfluid/private/load_sf.f:40-58)."""
    with open(os.path.join(GAPS, "%s_%s.data" % (name, code))) as f:
        for _ in range(3):
            f.readline()
        nv, kmin, kmax = f.readline().split()
        vals = f.read().split()
    nv = int(nv)
    kF = np.array(vals[:nv], dtype=float)
    Tc = np.array(vals[nv:2 * nv], dtype=float)
    return kF, Tc, float(kmin), float(kmax)
