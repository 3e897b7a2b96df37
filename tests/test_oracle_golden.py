"""Pin the CPU oracle against every data-free golden vector the reference's own tests hold
(SURVEY.md section 8c).  Copies of the reference's sample outputs live in tests/golden/reference/.
Tolerances are the reference's own ndiff tolerances (<module>/test/chk:15) unless the printed
precision is coarser."""
import os

import numpy as np
import pytest

from fixtures import GOLD, load_gap_file, read_rows

KERG = 1.380649e-16
MEV = 1.602176634e-6


def test_constants_golden(oracle):
    # constants/test/sample_output.data:3-51, ndiff --relerr 1e-6 (constants/test/chk:15); 14 digits printed
    names, vals = [], []
    for line in open(os.path.join(GOLD, "constants.sample_output.data")):
        if "=" in line and "dstar" not in line and "CODATA" not in line:
            k, v = line.split("=")
            names.append(k.strip()); vals.append(float(v))
    got = oracle.constants()
    order = ["pi", "twopi", "threepi", "fourpi", "onethird", "twothird", "fivethird", "seventhird",
             "threepisquare", "boltzmann", "avogadro", "clight", "clight2", "amu", "hbar", "finestructure",
             "electroncharge", "GMsun", "Gnewton", "Msun", "Mneutron", "Mproton", "Melectron", "Mmuon",
             "mev_to_ergs", "fm_to_cm", "ergs_to_mev", "cm_to_fm", "a_Bohr", "Rydberg", "Thomson", "hbarc_n",
             "mass_n", "length_n", "density_n", "pressure_n", "temperature_n", "Mn_n", "Mp_n", "Me_n", "mass_g",
             "potential_g", "length_g", "density_g", "pressure_g", "time_g"]
    assert names == order
    np.testing.assert_allclose(got[:len(order)], vals, rtol=2e-13)


def test_nucchem_golden(oracle):
    # nucchem/test/sample_output:4-8 (cmp exact, f8.3)
    comp, xsum, nch, idx, yion = oracle.moments([26, 28], [56, 56], [0.5, 0.5], mass_fracs=True)
    A, Z, Z2, Ye, Q = comp[0], comp[1], comp[3], comp[8], comp[10]
    assert ["%8.3f" % v for v in (A, Z, Z2, Ye, Q)] == ["  56.000", "  27.000", " 730.000", "   0.482", "   1.000"]


def _hz90_like(oracle, zz, aa, xn):
    comp, xsum, nch, idx, yion = oracle.moments([0, zz], [1, aa], [xn, (1.0 - xn) / aa], exclude_neutrons=True)
    comp[10] = 4.0
    return comp, yion


ZZ = dict(zip(range(2, 15), [2, 2, 2, 2, 26, 26, 26, 26, 24, 20, 14, 24, 24]))
AA = dict(zip(range(2, 15), [4, 4, 4, 4, 56, 56, 56, 56, 56, 56, 46, 96, 96]))
XN = dict(zip(range(2, 15), [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.07, 0.76, 0.80]))


def test_crust_neutrino_golden(oracle):
    # neutrino/test/sample_output.data:61-159, ndiff --relerr 1e-4 (neutrino/test/chk:15).
    # pair/photo/plasma/brems are data-free; pbf needs the 'gc' gap (rebuilt from the superfluid golden).
    rows = read_rows(os.path.join(GOLD, "neutrino.sample_output.data"), after="crust neutrino emissivity", ncol=11)
    assert len(rows) == 99
    kF, Tc, kmin, kmax = load_gap_file("gc", "n1")
    oracle.sf_load(1, kF, Tc, kmin, kmax)
    amu = 1.0 / 6.02214076e23
    worst = 0.0
    for r in rows:
        lgr, lgT = r[0], r[1]
        i = int(round(lgr))
        comp, _ = _hz90_like(oracle, ZZ[i], AA[i], XN[i])
        rho = 10.0 ** i
        T = np.exp(lgT * np.log(10.0))
        chi = (4.0 * np.pi / 3.0) * (float(np.float32(1.13)) * 1e-13) ** 3 * (rho * (1.0 - comp[9]) / amu)
        nn = rho * comp[9] / amu / (1.0 - chi)
        kn = (1.5 * np.pi ** 2 * nn) ** (1.0 / 3.0) / 1e13
        Tcn = oracle.sf_get_results(0.0, kn)[1]
        eps = oracle.crust_neutrino([rho], [T], [comp], [chi], [Tcn])[0]
        # columns: total pair photo plasma brems pbf
        for col, name in ((1, "pair"), (2, "photo"), (3, "plasma"), (4, "brems")):
            ref = r[5 + col]
            if ref == 0.0:
                assert eps[col] == 0.0 or eps[col] < 1e-290, (name, r, eps)
            elif ref > 1e-280:
                rel = abs(eps[col] / ref - 1.0)
                worst = max(worst, rel)
                assert rel < 1e-4, (name, lgr, lgT, ref, eps[col])
    assert worst < 1e-4


def test_conductivity_golden(oracle):
    # conductivity/test/sample_output.data:3-210 (208 rows), ndiff --relerr 1e-5 (conductivity/test/chk:15).
    # Gamma, eta_e come out of the oracle EOS (HELM table rebuilt from first principles), everything else
    # out of the oracle conductivity, neutrons + superfluid phonons on, Tc = 1e8 (test_cond.f:44-46,86).
    rows = read_rows(os.path.join(GOLD, "conductivity.sample_output.data"), skip=2, ncol=15)
    assert len(rows) == 208
    ctrl = [1, 1, 1, 1, 14.5, 14.5]
    tol = {"Gamma": 2e-6, "eta_e": 2e-6, "K_tot": 1e-5, "K_ee": 1e-5, "K_ei": 1e-5, "K_eQ": 1e-5,
           "K_sF": 1e-5, "K_nQ": 1e-5, "K_nph": 1e-5, "kappa": 1e-5}
    worst = {k: 0.0 for k in tol}
    for r in rows:
        i = int(round(r[0]))
        comp, yion = _hz90_like(oracle, ZZ[i], AA[i], XN[i])
        rho = 10.0 ** i
        T = 10.0 ** r[1]
        ierr, res, phase, chi = oracle.eval_crust_eos([rho], [T], [comp], [ZZ[i]], [AA[i]], [yion], [[1e8] * 3])
        assert ierr == 0
        res = res[0]
        Gamma, Theta, mu_e = res[9], res[10], res[11]
        K = oracle.conductivity([rho], [T], chi, [Gamma], [Theta], [mu_e], [comp], [1e8], ctrl)[0]
        got = {"Gamma": Gamma, "eta_e": mu_e * MEV / KERG / T, "K_tot": K[0], "K_ee": K[1], "K_ei": K[2],
               "K_eQ": K[3], "K_sF": K[4], "K_nQ": K[5], "K_nph": K[6], "kappa": K[7]}
        ref = dict(zip(["Gamma", "eta_e", "K_tot", "K_ee", "K_ei", "K_eQ", "K_sF", "K_nQ", "K_nph", "kappa"], r[5:]))
        for k in tol:
            if ref[k] == 0.0:
                assert got[k] == 0.0, (k, r[:2], got[k])
            else:
                rel = abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-300)
                if k == "eta_e":
                    rel = abs(got[k] - ref[k]) / max(abs(ref[k]), 1.0)
                worst[k] = max(worst[k], rel)
                assert rel < tol[k], (k, r[:2], ref[k], got[k], rel)
    print("conductivity golden worst rel:", worst)


def test_pcy97_atm_golden(oracle):
    # dStar_atm/test/sample_output.data pcy97 blocks, ndiff --relerr 1e-3 (dStar_atm/test/chk:15), 5 decimals.
    # Pins pcy97.f, interp_pm and interp_value_and_slope (value AND slope, incl. the end knot).
    txt = open(os.path.join(GOLD, "dStar_atm.sample_output.data")).read().split("\n")
    blocks, cur = [], None
    for line in txt:
        if line.strip() in ("pcy97", "bc09"):
            cur = {"kind": line.strip(), "rows": []}
            blocks.append(cur)
        elif cur is not None and line.startswith("Plight"):
            cur["lgPlight"] = float(line.split("=")[1])
        elif cur is not None and len(line.split()) == 5:
            try:
                cur["rows"].append([float(v) for v in line.split()])
            except ValueError:
                pass
    pcy = [b for b in blocks if b["kind"] == "pcy97"]
    assert len(pcy) == 2 and all(len(b["rows"]) == 20 for b in pcy)
    g = 2.43e14
    for b, Plight in zip(pcy, (2.43e19, 2.43e23)):
        lgTb_tab = 7.0 + np.arange(256) * (9.5 - 7.0) / 255.0
        lgTeff, lgflux = oracle.pcy97(g, Plight, lgTb_tab)
        fT = oracle.interp_pm(lgTb_tab, lgTeff)
        fF = oracle.interp_pm(lgTb_tab, lgflux)
        rows = np.array(b["rows"])
        x = 7.0 + 1.5 * np.arange(20) / 19.0
        vT, sT = oracle.interp_value_and_slope(lgTb_tab, fT, x)
        vF, sF = oracle.interp_value_and_slope(lgTb_tab, fF, x)
        np.testing.assert_allclose(vT, rows[:, 1], atol=6e-6)
        np.testing.assert_allclose(sT, rows[:, 2], atol=6e-6)
        np.testing.assert_allclose(vF, rows[:, 3], atol=6e-6)
        np.testing.assert_allclose(sF, rows[:, 4], atol=2e-5)


def test_superfluid_golden(oracle):
    # superfluid/test/sample_output.data (ndiff --relerr 1e-6): the three rebuilt tables reproduce the
    # golden rows at their own knots; pins sf_get_Tc's range clipping and interp_value.
    rows = read_rows(os.path.join(GOLD, "superfluid.sample_output.data"), skip=2, ncol=4)
    for typ, (name, code) in enumerate((("ccdk93", "p1"), ("gc", "n1"), ("ao85", "n3"))):
        kF, Tc, kmin, kmax = load_gap_file(name, code)
        oracle.sf_load(typ, kF, Tc, kmin, kmax)
    oracle.sf_scale([1.0, 1.0, 1.0])
    for i, r in enumerate(rows):
        k = 3.6 * i / 39.0
        tc = oracle.sf_get_results(k, k) * 1e-9
        np.testing.assert_allclose(tc, r[1:], atol=1.1e-6)


def test_eos_golden(oracle):
    # dStar_eos/test/sample_output.data (9 densities x 11 temperatures x 13 columns), ndiff --relerr 1e-4
    # (dStar_eos/test/chk:15).  The HELM table is rebuilt from first principles (tools/gen_helm_table.cpp),
    # the 'gc' n1S0 gap from the superfluid golden; printed precision is 4-5 significant digits / 3 decimals.
    txt = open(os.path.join(GOLD, "dStar_eos.sample_output.data")).read().split("\n")
    kF, Tc, kmin, kmax = load_gap_file("gc", "n1")
    oracle.sf_load(1, kF, Tc, kmin, kmax)
    oracle.sf_load(0, [0.0, 1.0], [0.0, 0.0], 0.0, 0.0)
    oracle.sf_load(2, [0.0, 1.0], [0.0, 0.0], 0.0, 0.0)
    oracle.sf_scale([1.0, 1.0, 1.0])
    mixes = {0: [0.0, 0.5, 0.5], 1: [0.5, 0.0, 0.5], 2: [0.95, 0.0, 0.05]}
    Z = np.array([0.0, 8.0, 26.0]); A = np.array([1.0, 16.0, 56.0])
    block = -1
    amu = 1.0 / 6.02214076e23
    nrows = 0
    worst = np.zeros((2, 13))
    for line in txt:
        if line.startswith("<Z>"):
            block += 1
            comp, xsum, nch, idx, yion = oracle.moments(Z, A, np.array(mixes[block]) / A, exclude_neutrons=True)
            continue
        p = line.split()
        if len(p) != 15:
            continue
        try:
            r = [float(v) for v in p]
        except ValueError:
            continue
        rho = 10.0 ** r[0]
        i = int(round((r[1] - 7.5) * 10))
        lgT = float(np.float32(7.5) + np.float32(i) / np.float32(10.0))  # REAL(4) arithmetic in test_crust_eos.f:96
        T = 10.0 ** lgT
        chi = (4.0 * np.pi / 3.0) * (float(np.float32(1.13)) * 1e-13) ** 3 * (rho * (1.0 - comp[9]) / amu)
        n = rho * comp[9] / amu / (1.0 - chi)
        kn = (3.0 * np.pi ** 2 * n) ** (1.0 / 3.0) / 1e13
        Tcs = oracle.sf_get_results(0.0, kn)
        ierr, res, phase, chi_o = oracle.eval_crust_eos([rho], [T], [comp], Z[1:], A[1:], [yion], [Tcs], chi_in=[chi])
        assert ierr == 0
        res = res[0]
        ln10 = np.log(10.0)
        got = [res[9], res[10], res[7] / KERG / 6.02214076e23, res[1] / ln10, res[0] / ln10, res[2] / ln10,
               res[4], res[5], res[13], res[14], res[3], res[11], res[12]]
        ref = r[2:]
        nrows += 1
        # Columns that feel the n 1S0 critical temperature (neutron Cv and S carry the superfluid
        # reduction R(T/Tc), neutron.f:57-66): Cv, lg S, G3-1, del_ad.  The 'gc' table is resampled
        # from 40 golden points, so in the two neutron-rich mixes these are only checked loosely.
        gap_cols = (2, 5, 9, 10)
        for c in range(13):
            loose = block >= 1 and c in gap_cols
            if c < 3:   # es12.4 columns: 5 significant digits
                err = abs(got[c] - ref[c]) / abs(ref[c])
                assert err < (5e-3 if loose else 1.2e-4), (c, r[:2], ref[c], got[c])
            else:       # f8.3 columns
                err = abs(got[c] - ref[c])
                assert err < (2e-3 if loose else 1.5e-3), (c, r[:2], ref[c], got[c])
            worst[min(block, 1), c] = max(worst[min(block, 1), c], err)
    assert nrows == 99
    print("eos golden worst (O-Fe mix; neutron-rich mixes):", worst)
