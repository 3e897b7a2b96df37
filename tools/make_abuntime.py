"""Synthetic "abuntime" network-output file (format: reference dStar_crust/preprocessor/README:3-31) for
BASELINE.json configs[4] (INT-16-2b-demo flavour): the real rp-process ash files are absent (500 MB, not in the
reference tree), so this writes a file of the same format with 63 nuclides over Z = 8..40 plus free neutrons,
Dirichlet(alpha = 0.3) mass fractions drawn at anchor depths every 0.5 dex in lg P and interpolated in between,
from lg P = 26 to 30.6 in steps of 0.002 dex (finer than the preprocessor's 0.005 cull, so the cull is exercised).
Seed 20261019.  Usage: python tools/make_abuntime.py OUT [n_steps]
"""
import sys

import numpy as np

SEED = 20261019
GRAVITY, MDOT = 1.85052e14, 8.8e4 * 0.3          # abuntime.f:9-10
EL = ["n", "h", "he", "li", "be", "b", "c", "n", "o", "f", "ne", "na", "mg", "al", "si", "p", "s", "cl", "ar", "k", "ca",
      "sc", "ti", "v", "cr", "mn", "fe", "co", "ni", "cu", "zn", "ga", "ge", "as", "se", "br", "kr", "rb", "sr", "y", "zr"]


def network():
    names, Z, A = ["n"], [0], [1]
    for z in range(8, 41):
        for a in (int(round(2.15 * z)), int(round(2.45 * z))):
            if len(names) < 64:
                names.append("%s%d" % (EL[z], a)); Z.append(z); A.append(a)
    return names, np.array(Z, float), np.array(A, float)


def table(lgP, names, A, rng):
    anchors = np.arange(26.0, 31.01, 0.5)
    Xa = rng.dirichlet(np.full(len(names) - 1, 0.3), size=len(anchors))          # charged species, mass fractions
    X = np.empty((len(lgP), len(names)))
    for j in range(len(names) - 1):
        X[:, j + 1] = np.interp(lgP, anchors, Xa[:, j])
    X[:, 1:] /= X[:, 1:].sum(axis=1, keepdims=True)
    Xn = np.clip((lgP - 29.5) / (30.6 - 29.5), 0.0, 1.0) * 0.3                   # neutron drip sets in at lgP ~ 29.5
    X[:, 1:] *= (1.0 - Xn)[:, None]
    X[:, 0] = Xn
    return X / A[None, :]                                                         # abundances Y = X / A


def write(path, n_steps=2301):
    rng = np.random.default_rng(SEED)
    names, Z, A = network()
    lgP = np.linspace(26.0, 30.6, n_steps)
    Y = table(lgP, names, A, rng)
    with open(path, "w") as f:
        f.write(" %5d\n" % len(names))
        for i in range(n_steps):
            time = 10.0 ** lgP[i] / (GRAVITY * MDOT)
            f.write(" %18.10E%11.3E%11.3E  %18.10E%18.10E\n" % (time, 0.5, 1.0e9, 1.0, 0.0))
            for k0 in range(0, len(names), 5):
                f.write(" " + "".join("%5s%10.3E" % (names[k], Y[i, k]) for k in range(k0, min(k0 + 5, len(names)))) + "\n")
            f.write("0" + "-" * 79 + "\n")
    return names, Z, A, lgP, Y


if __name__ == "__main__":
    write(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2301)
