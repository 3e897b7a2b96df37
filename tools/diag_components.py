"""Diagnostic (GPU box): per-component EOS deviations between the CUDA path and the oracle."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import dstar_b200 as ds
from dstar_b200._capi import p, ip, f64
from oracle_lib import Oracle, _ptr
from parity_helpers import load_oracle_gaps
from test_gpu_parity import NET_A, NET_Z, R_NUC, AMU, random_crust_points

o = Oracle(); load_oracle_gaps(o, "ns", "gc", "t72")
ctx = ds.Context(0, gaps=("ns", "gc", "t72"))
rng = np.random.default_rng(20261019)
n = 4000
rho, T, ionic, Yion = random_crust_points(o, n, rng)
chi = (4 * np.pi / 3) * R_NUC ** 3 * (rho * (1 - ionic[:, 9]) / AMU)
kn = (3 * np.pi ** 2 * rho * ionic[:, 9] / AMU / (1 - chi)) ** (1 / 3) / 1e13
Tcs = np.array([o.sf_get_results(0.0, k) for k in kn])
res_o = np.zeros((n, 15)); ph = np.zeros(n, dtype=np.int32); chio = np.zeros(n); co = np.zeros((n, 4, 7))
a = [f64(x) for x in (rho, T, ionic, NET_Z, NET_A, Yion, Tcs)]
o.lib.dso_eval_crust_eos_c(n, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), len(NET_Z), _ptr(a[3]), _ptr(a[4]), _ptr(a[5]),
                           _ptr(a[6]), None, None, _ptr(res_o), ph.ctypes.data_as(ip), _ptr(chio), _ptr(co))
res_g = np.zeros((n, 15)); cg = np.zeros((n, 4, 7)); ie = np.zeros(n, dtype=np.int32)
ctx.L.dstar_eval_crust_eos_components(ctx.h, n, p(a[0]), p(a[1]), p(a[2]), len(NET_Z), p(a[3]), p(a[4]), p(a[5]),
                                      p(a[6]), None, None, p(res_g), p(ph, ip), p(chio), p(ie, ip), p(cg))
names = ["P", "E", "S", "F", "Cv", "dPdlnRho", "dPdlnT"]
for q, comp in enumerate(["electron", "ion", "neutron", "radiation"]):
    for k, nm in enumerate(names):
        r = co[:, q, k]; g = cg[:, q, k]
        m = np.abs(r) > 0
        if not m.any():
            continue
        d = np.abs(g[m] - r[m]) / np.abs(r[m])
        i = np.argmax(d)
        idx = np.where(m)[0][i]
        print("%-9s %-9s max rel %.2e  at rho=%.3e T=%.3e Yn=%.3f  frac>1e-10: %.4f" %
              (comp, nm, d[i], rho[idx], T[idx], ionic[idx, 9], np.mean(d > 1e-10)))
