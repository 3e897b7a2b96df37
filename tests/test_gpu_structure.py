"""create_model's structure arrays against independent restatements:

* the crust table lg rho(lg P), lg eps(lg P) of do_generate_crust_table / find_densities
  (dStar_crust/private/dStar_crust_mod.f:74-171, :282-368; match_density :370-443): at every table point the ORACLE's
  EOS, evaluated at the tabulated density, must return the tabulated pressure (the root find's own tolerance is 1e-12
  in lg P), and lg eps must be rho (1 + sum Y ME / amu + E / c^2) with the oracle's internal energy;
* the TOV integration and zone geometry (NScool/private/NScool_crust_tov.f:110-303 with dense output every
  target_resolution_lnP, NScool_create_model.f:138-243), restated here in numpy on top of scipy's DOP853 at a tolerance
  1e6 times tighter than the reference's (rtol 1e-5, atol 1e-6): the product's arrays must agree to the reference's
  integration tolerance, the zone count and pressure grid exactly."""
import numpy as np
import pytest

from parity_helpers import load_oracle_gaps

pytestmark = pytest.mark.gpu

# constants_def.f:40-68 (CODATA 2018 as the product and the oracle use them)
G = 6.67430e-8; C2 = 2.99792458e10 ** 2; MSUN = 1.3271244e26 / G
LEN_G = G * MSUN / C2; DEN_G = MSUN / LEN_G ** 3; P_G = DEN_G * C2
LN10 = np.log(10.0)
AMU = 1.0 / 6.02214076e23


@pytest.fixture(scope="module")
def nscool():
    import dstar_b200 as ds
    ds.NScool_init()
    yield ds
    ds.NScool_shutdown()


def cubic(x, coef, xv):
    """interp_value on (4, n) monotone-cubic coefficients (MESA interp_1d)."""
    k = np.clip(np.searchsorted(x, xv, side="right") - 1, 0, len(x) - 2)
    dx = xv - x[k]
    c = coef.reshape(-1, 4)[k]
    return c[:, 0] + dx * (c[:, 1] + dx * (c[:, 2] + dx * c[:, 3])) if np.ndim(xv) else c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]))


def test_crust_table_inverts_the_oracle_eos(nscool, oracle):
    ds = nscool
    s = ds.NScool(None, target_resolution_lnP=0.5, which_neutron_1S0_gap="sfb03")
    s.create_model()
    lgP = s.array("crust_lgP"); n = len(lgP)
    lgRho = s.array("crust_lgRho").reshape(n, 4)[:, 0]; lgEps = s.array("crust_lgEps").reshape(n, 4)[:, 0]
    nch = s.get_int("ncharged")
    ionic = s.array("crust_ionic").reshape(n, 11); Yion = s.array("crust_Yion").reshape(n, nch)
    Tref = s.array("crust_Tref")[0]
    ME = s.array("crust_mass_excess")
    assert n == 2048 and Tref == 1.0e8                              # generate_HZ90.f:12-14, dStar_crust_def.f default
    pick = np.arange(0, n, 7)                                        # every 7th point: all HZ90 layers, both drips
    rho = 10.0 ** lgRho[pick]
    Yn = ionic[pick, 9]
    chi = (4.0 * np.pi / 3.0) * (float(np.float32(1.13)) * 1.0e-13) ** 3 * (np.minimum(rho, 1.0e14) * (1.0 - Yn) / AMU)
    kn = (3.0 * np.pi ** 2 * rho * Yn / AMU / (1.0 - chi)) ** (1.0 / 3.0) / 1.0e13
    load_oracle_gaps(oracle, "ns", "sfb03", "t72")
    Tcs = np.array([oracle.sf_get_results(0.0, k) for k in kn])
    ie, res, ph, _ = oracle.eval_crust_eos(rho, np.full(len(pick), Tref), ionic[pick], s.array("Zion"), s.array("Aion"),
                                            Yion[pick], Tcs, chi_in=chi)
    assert ie == 0
    dev = np.abs(res[:, 0] / LN10 - lgP[pick])
    print("crust table: max |lg P_oracle(rho_tab) - lg P_tab| =", dev.max())
    assert dev.max() < 1e-10
    eps = rho * (1.0 + Yion[pick] @ ME / 931.49410242 + np.exp(res[:, 1]) / C2)
    assert np.max(np.abs(np.log10(eps) - lgEps[pick])) < 1e-9
    assert np.all(np.diff(lgRho) > 0)
    s.free()


@pytest.mark.parametrize("rez,Mcore,Rcore", [(0.2, 1.6, 11.0), (0.05, 1.4, 10.0)])
def test_structure_against_independent_tov(nscool, rez, Mcore, Rcore):
    from scipy.integrate import solve_ivp
    ds = nscool
    s = ds.NScool(None, target_resolution_lnP=rez, core_mass=Mcore, core_radius=Rcore, which_neutron_1S0_gap="sfb03",
                  core_temperature=1.0e8)
    s.create_model()
    nz = s.get_int("nz")
    lgPt = s.array("crust_lgP"); cR = s.array("crust_lgRho"); cE = s.array("crust_lgEps")
    Rc = Rcore * 1.0e5 / LEN_G

    def rhs(lnP, y):   # tov_derivs_crust, NScool_crust_tov.f:191-236
        P = np.exp(lnP)
        r = y[0] + Rc; m = y[2] + Mcore
        lgP = np.clip(lnP / LN10 + np.log10(P_G), lgPt[0], lgPt[-1])
        rho_g = 10.0 ** cubic(lgPt, cR, lgP) / DEN_G; eps_g = 10.0 ** cubic(lgPt, cE, lgP) / DEN_G
        Lam = 1.0 / np.sqrt(1.0 - 2.0 * m / r); H = 1.0 + P / eps_g; Gf = 1.0 + 4.0 * np.pi * r ** 3 * P / m
        g = m / r ** 2 * Gf * Lam
        d0 = -P / g / eps_g / H / Lam
        return [d0, -P * 4.0 * np.pi * r * r / g / H * rho_g / eps_g, -P * 4.0 * np.pi * r * r / g / H / Lam, -P / eps_g / H,
                4.0 * np.pi * r * r * Lam * d0]

    lnP0 = 32.5 * LN10 - np.log(P_G); lnP1 = 27.0 * LN10 - np.log(P_G)
    y0 = [0.0, 0.0, 0.0, 0.5 * np.log(1.0 - 2.0 * Mcore / Rc), 0.0]
    sol = solve_ivp(rhs, (lnP0, lnP1), y0, method="DOP853", rtol=1e-11, atol=1e-12, dense_output=True)
    assert sol.success
    # what the reference's own tolerance (rtol 1e-5, atol 1e-6, NScool_crust_tov.f:143-144) buys with this integrator
    loose = solve_ivp(rhs, (lnP0, lnP1), y0, method="DOP853", rtol=1e-5, atol=1e-6, first_step=0.1)
    print("steps at the reference tolerance:", len(loose.t) - 1, "end-point deviation from the tight run:",
          np.abs(loose.y[:, -1] - sol.y[:, -1]) / np.abs(sol.y[:, -1]))
    # recorded points: lnP0 - j*rez while > lnP1 (tov_solout_crust :238-303)
    xs = []
    x = lnP0
    while x > lnP1:
        xs.append(x); x -= rez
    xs = np.array(xs)
    Y = sol.sol(xs)
    nzs = len(xs)
    assert nz == nzs - 1
    yend = sol.y[:, -1]
    phi = Y[3] + (0.5 * np.log(1.0 - 2.0 * (yend[2] + Mcore) / (yend[0] + Rc)) - yend[3])
    t = np.arange(nzs - 1, 0, -1)                                      # stov%x(nzs:2:-1)
    P_bar = np.exp(xs[t]) * P_G
    ePhi_bar = np.exp(phi[t])
    m = Y[1][t] * MSUN
    area = 4.0 * np.pi * (Y[0][t] * LEN_G + Rcore * 1.0e5) ** 2
    eLam_bar = 1.0 / np.sqrt(1.0 - 2.0 * (Y[2][t] + Mcore) / (Y[0][t] + Rc))
    dm = np.append(m[:-1] - m[1:], m[-1])
    rho = dm / (Y[4][t] - Y[4][t - 1]) / LEN_G ** 3
    ePhi = np.append(0.5 * (ePhi_bar[:-1] + ePhi_bar[1:]), 0.5 * (ePhi_bar[-1] + np.exp(phi[0])))
    Pz = np.append(0.5 * (P_bar[:-1] + P_bar[1:]), 1.5 * P_bar[-1] - 0.5 * P_bar[-2])
    dm_bar = np.append(0.5 * dm[0], 0.5 * (dm[:-1] + dm[1:]))
    rho_bar = 0.5 * (rho[:-1] * dm[1:] + rho[1:] * dm[:-1]) / dm_bar[1:]
    T = 1.0e8 * np.exp(phi[0]) / ePhi

    def rel(a, b):
        return np.max(np.abs(a / b - 1.0))
    assert rel(s.array("P_bar"), P_bar) < 1e-12 and rel(s.array("P"), Pz) < 1e-12        # the grid is exact
    # The reference integrates at rtol 1e-5 / atol 1e-6 and takes the zone values from the dense output: with scipy's
    # DOP853 at that same tolerance the end point is 7e-4 off the converged one in radius and volume (printed above),
    # and a zone's density -- baryon mass over a DIFFERENCE of two interpolated volumes -- is noisier still.  The
    # bounds below are that tolerance-limited accuracy (measured: metric 1e-5, area 7e-5, dm 5e-4, rho 2e-2), not a
    # claim of agreement with the Fortran path's bits, which nothing here can check.
    tol = 1.0e-4
    loose_tol = {"dm": 1.0e-3, "dm_bar": 1.0e-3, "rho": 5.0e-2}
    worst = {}
    for name, want in (("ePhi_bar", ePhi_bar), ("ePhi", ePhi), ("area", area), ("eLambda_bar", eLam_bar), ("dm", dm),
                       ("rho", rho), ("dm_bar", dm_bar), ("T", T)):
        worst[name] = rel(s.array(name), want)
    print("nz =", nz, "worst relative deviations from the tight integration:", {k: "%.1e" % v for k, v in worst.items()})
    for name in worst:
        assert worst[name] < loose_tol.get(name, tol), (name, worst[name])
    assert rel(s.array("rho_bar")[1:], rho_bar) < loose_tol["rho"]
    assert abs(np.sum(s.array("dm")) / np.sum(dm) - 1.0) < tol           # the crust's baryon mass as a whole
    assert rel(s.array("e2Phi"), ePhi ** 2) < 2 * tol and rel(s.array("lnT"), np.log(T)) < tol
    Mtot = yend[2] + Mcore; Rtot = yend[0] * LEN_G * 1.0e-5 + Rcore
    assert abs(s.get_real("Mtotal") / Mtot - 1.0) < tol and abs(s.get_real("Rtotal") / Rtot - 1.0) < tol
    grav = Mtot * MSUN * G / (Rtot * 1.0e5) ** 2 * eLam_bar[0]
    assert abs(s.get_real("grav") / grav - 1.0) < 3 * tol              # ~ 2 dR/R + d eLambda
    s.free()
