"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: 1e-10 relative per microphysics coefficient,
1e-6 relative on Teff(t)."""
import numpy as np
import pytest

from parity_helpers import load_oracle_gaps, oracle_evolve_star, oracle_micro_tables

pytestmark = pytest.mark.gpu

SEED = 20261019
AMU = 1.0 / 6.02214076e23
R_NUC = float(np.float32(1.13)) * 1e-13

# HZ90 layers (dStar_crust/preprocessor/src/HZ90_comp.f:29-44): (Z, A, Xn, lg rho at which it is typical)
LAYERS = [(26, 56, 0.0, 8.0), (24, 56, 0.0, 9.5), (22, 56, 0.0, 10.3), (20, 56, 0.0, 10.8), (18, 56, 0.0, 11.2),
          (16, 52, 4 / 56, 11.6), (14, 46, 10 / 56, 11.9), (12, 40, 16 / 56, 12.1), (20, 68, 22 / 56, 12.3),
          (18, 62, 50 / 112, 12.5), (16, 56, 56 / 112, 12.7), (14, 50, 62 / 112, 12.9), (12, 44, 68 / 112, 13.0),
          (18, 66, 158 / 224, 13.2), (16, 60, 164 / 224, 13.4), (14, 54, 170 / 224, 13.6), (12, 48, 176 / 224, 13.8),
          (22, 88, 360 / 448, 14.0)]
NET_Z = np.array([12, 12, 12, 14, 14, 14, 16, 16, 16, 18, 18, 18, 20, 20, 22, 22, 24, 26], dtype=float)
NET_A = np.array([40, 44, 48, 46, 50, 54, 52, 56, 60, 56, 62, 66, 56, 68, 56, 88, 56, 56], dtype=float)


@pytest.fixture(scope="module")
def ctx():
    import dstar_b200 as ds
    c = ds.Context(0, gaps=("ns", "gc", "t72"))
    yield c
    c.close()


def random_crust_points(oracle, n, rng, mix=True):
    """n random (rho, T, composition) points along an HZ90-like crust; two-species mixes near transitions."""
    rho = np.zeros(n); T = 10.0 ** rng.uniform(7.0, 10.0, n)
    ionic = np.zeros((n, 11)); Yion = np.zeros((n, len(NET_Z)))
    for i in range(n):
        l = rng.integers(0, len(LAYERS))
        Z, A, Xn, lgr = LAYERS[l]
        rho[i] = 10.0 ** (lgr + rng.uniform(-0.3, 0.3))
        Y = np.zeros(len(NET_Z) + 1)  # neutron first
        Zs = np.concatenate([[0.0], NET_Z]); As = np.concatenate([[1.0], NET_A])
        j = 1 + int(np.where((NET_Z == Z) & (NET_A == A))[0][0])
        if mix and rng.random() < 0.3 and l + 1 < len(LAYERS):
            Z2, A2, Xn2, _ = LAYERS[l + 1]
            j2 = 1 + int(np.where((NET_Z == Z2) & (NET_A == A2))[0][0])
            w = rng.uniform(0.05, 0.95)
            Y[0] = w * Xn + (1 - w) * Xn2
            Y[j] = w * (1 - Xn) / A
            Y[j2] = (1 - w) * (1 - Xn2) / A2
        else:
            Y[0] = Xn; Y[j] = (1 - Xn) / A
        comp, xsum, nch, idx, yion = oracle.moments(Zs, As, Y, exclude_neutrons=True)
        comp[10] = rng.uniform(0.0, 30.0)
        ionic[i] = comp; Yion[i] = yion
    return rho, T, ionic, Yion


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    den = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / den)


def test_device_present(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10, info   # the library is built for sm_100a only
    tf = ctx.fp64_peak_tflops()
    assert tf > 1.0, tf


def test_eos_parity(ctx, oracle):
    rng = np.random.default_rng(SEED)
    n = 4000
    rho, T, ionic, Yion = random_crust_points(oracle, n, rng)
    load_oracle_gaps(oracle, "ns", "gc", "t72")
    chi = (4 * np.pi / 3) * R_NUC ** 3 * (rho * (1 - ionic[:, 9]) / AMU)
    kn = (3 * np.pi ** 2 * rho * ionic[:, 9] / AMU / (1 - chi)) ** (1 / 3) / 1e13
    Tcs = np.array([oracle.sf_get_results(0.0, k) for k in kn])
    Tcs_gpu = ctx.sf_get_results(np.zeros(n), kn)
    assert rel(Tcs_gpu[:, 1], Tcs[:, 1]) < 1e-13
    ie_o, res_o, ph_o, chi_o = oracle.eval_crust_eos(rho, T, ionic, NET_Z, NET_A, Yion, Tcs)
    ie_g, res_g, ph_g, chi_g = ctx.eval_crust_eos(rho, T, ionic, NET_Z, NET_A, Yion, Tcs)
    assert ie_o == 0 and not ie_g.any()
    assert np.array_equal(ph_o, ph_g)
    assert rel(chi_g, chi_o) < 1e-13 and rel(chi_g, chi) < 1e-12   # nuclear volume fraction (dStar_eos_lib.f:123 out arg)
    names = ["lnP", "lnE", "lnS", "grad_ad", "chiRho", "chiT", "Cp", "Cv", "Chi", "Gamma", "Theta", "mu_e", "mu_n",
             "Gamma1", "Gamma3"]
    worst = {}
    for k, nm in enumerate(names):
        if nm in ("lnP", "lnE", "lnS"):   # logs: compare the quantity itself
            worst[nm] = float(np.max(np.abs(res_g[:, k] - res_o[:, k])))
        else:
            worst[nm] = float(rel(res_g[:, k], res_o[:, k]))
    print("EOS parity worst:", worst)
    for nm, w in worst.items():
        assert w < 1e-10, (nm, w)


def test_neutrino_parity(ctx, oracle):
    rng = np.random.default_rng(SEED + 1)
    n = 4000
    rho, T, ionic, _ = random_crust_points(oracle, n, rng)
    chi = (4 * np.pi / 3) * R_NUC ** 3 * (rho * (1 - ionic[:, 9]) / AMU)
    Tcn = 10.0 ** rng.uniform(7.5, 10.0, n)
    eo = oracle.crust_neutrino(rho, T, ionic, chi, Tcn)
    eg = ctx.crust_neutrino(rho, T, ionic, chi, Tcn)
    for c, nm in enumerate(["total", "pair", "photo", "plasma", "brems", "pbf"]):
        m = np.abs(eo[:, c]) > 1e-250
        w = rel(eg[m, c], eo[m, c])
        print("neutrino", nm, w)
        assert w < 1e-10, (nm, w)
        assert np.all(eg[~m, c] < 1e-240)


@pytest.mark.parametrize("neutrons", [False, True])
def test_conductivity_parity(ctx, oracle, neutrons):
    rng = np.random.default_rng(SEED + 2)
    n = 1500 if neutrons else 4000
    rho, T, ionic, Yion = random_crust_points(oracle, n, rng)
    if neutrons:
        T = 10.0 ** rng.uniform(7.0, 9.0, n)
    chi = (4 * np.pi / 3) * R_NUC ** 3 * (rho * (1 - ionic[:, 9]) / AMU)
    Tcs = np.tile([0.0, 3e9, 0.0], (n, 1))
    _, res, _, _ = oracle.eval_crust_eos(rho, T, ionic, NET_Z, NET_A, Yion, Tcs)
    ctrl = [1, 1, 1, 1, 10.0, 9.0] if neutrons else [0, 0, 1, 1, 10.0, 9.0]
    Ko = oracle.conductivity(rho, T, chi, res[:, 9], res[:, 10], res[:, 11], ionic, Tcs[:, 1], ctrl)
    Kg = ctx.conductivity(rho, T, chi, res[:, 9], res[:, 10], res[:, 11], ionic, Tcs[:, 1], ctrl)
    for c, nm in enumerate(["total", "ee", "ei", "eQ", "sf", "nQ", "np", "kap", "electron_total", "neutron_total"]):
        m = Ko[:, c] != 0.0
        w = rel(Kg[m, c], Ko[m, c]) if m.any() else 0.0
        print("conductivity", nm, w)
        assert w < 1e-10, (nm, w)
        assert np.all(Kg[~m, c] == 0.0)
    # the other scattering formulae
    ctrl2 = [0, 0, 2, 2, 10.0, 9.0]
    Ko = oracle.conductivity(rho[:500], T[:500], chi[:500], res[:500, 9], res[:500, 10], res[:500, 11], ionic[:500],
                             Tcs[:500, 1], ctrl2)
    Kg = ctx.conductivity(rho[:500], T[:500], chi[:500], res[:500, 9], res[:500, 10], res[:500, 11], ionic[:500],
                          Tcs[:500, 1], ctrl2)
    assert rel(Kg[:, 0], Ko[:, 0]) < 1e-10


def _make_star(ds, **kw):
    base = dict(target_resolution_lnP=0.2, number_epochs=2, basic_epoch_Mdots=[1.0e17, 0.0],
                basic_epoch_boundaries=[0.0, 50.0, 1050.0], core_temperature=1.4e8,
                atmosphere_temperature_when_accreting=2.4e8, turn_on_extra_heating=True, Q_heating_shallow=0.1,
                lgP_min_heating_shallow=27.5, lgP_max_heating_shallow=28.5, which_neutron_1S0_gap="sfb03",
                fix_Qimp=True, Qimp=1.0)
    base.update(kw)
    return ds.NScool(None, **base)


@pytest.fixture(scope="module")
def nscool():
    import dstar_b200 as ds
    ds.NScool_init()
    yield ds
    ds.NScool_shutdown()


def _check_tables(got_star, ref, tol=1e-10):
    worst = {}
    for name in ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK"):
        g = got_star.array(name).reshape(-1, 4)
        r = ref[name].reshape(-1, 4)
        # f(1,:) are logs of the coefficients: |d ln x| is the relative deviation of x itself
        worst[name] = float(np.max(np.abs(g[:, 0] - r[:, 0])))
        assert worst[name] < tol, (name, worst[name])
        # spline coefficients: same formulas on nearly identical knot values
        scale = np.maximum(np.abs(r[:, 1:]), 1e-3)
        assert np.max(np.abs(g[:, 1:] - r[:, 1:]) / scale) < 1e-6, name
    return worst


def test_micro_tables_parity_basic(nscool, oracle):
    """configs[0] (examples/basic_run inlist) at reduced resolution: every (4*128, nz) table entry."""
    s = _make_star(nscool)
    s.create_model()
    ref = oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72"))
    assert ref["ierr"] == 0
    worst = _check_tables(s, ref)
    print("micro tables worst |d ln|:", worst, "nz =", s.get_int("nz"))
    assert abs(s.array("rho_bar")[0] / ref["rho_bar1"] - 1.0) < 1e-12
    np.testing.assert_array_equal(s.array("tab_lnT"), ref["tab_lnT"])
    s.free()


def test_micro_tables_parity_neutron_channels(nscool, oracle):
    """configs[3] flavour: neutron + superfluid-phonon conductivity and shell Urca on (two DOP853 quadratures
    per face and knot on the device)."""
    s = _make_star(nscool, target_resolution_lnP=0.6, use_neutron_conductivity=True,
                   use_superfluid_phonon_conductivity=True, turn_on_shell_Urca=True, core_temperature=3.0e8,
                   atmosphere_temperature_when_accreting=5.0e8, which_neutron_1S0_gap="gc")
    s.create_model()
    ref = oracle_micro_tables(oracle, s, gaps=("ns", "gc", "t72"), cond_ctrl=(1, 1, 1, 1, 10.0, 9.0),
                              shell_Urca=(1, 1.0, 28.5))
    worst = _check_tables(s, ref)
    print("micro tables (neutron channels) worst |d ln|:", worst, "nz =", s.get_int("nz"))
    s.free()


def test_evolve_parity_basic(nscool, oracle):
    """Outburst (50 d at 1e17 g/s) + 1000 d of cooling: monitors, final profile and solver statistics."""
    s = _make_star(nscool)
    s.create_model()
    lnT0 = s.array("lnT").copy()
    s.evolve_model()
    t, Teff, Qb = s.monitors()
    ref = oracle_evolve_star(oracle, s, lnT0, heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, 0.1],
                             T_atm=2.4e8, ctrl=dict(turn_on_extra_heating=1))
    print("Teff_monitor gpu", Teff, "oracle", ref["Teff_monitor"], "counters", s.counters(), ref["counters"])
    assert rel(Teff, ref["Teff_monitor"]) < 1e-6
    assert rel(t, ref["t_monitor"]) < 1e-12
    assert rel(Qb[:1], ref["Qb_monitor"][:1]) < 1e-6 and Qb[1] == 0.0
    assert np.max(np.abs(s.array("lnT") - ref["lnT"])) < 1e-6
    np.testing.assert_array_equal(s.counters(), ref["counters"])
    s.free()


def test_evolve_batch_matches_single_and_oracle(nscool, oracle):
    """A small (Tc, Q_heating) sweep evolved as ONE batch: each star equals its own oracle run."""
    ds = nscool
    stars = []
    for Tc, Q in ((5e7, 0.0), (9.2e7, 1.7), (2e8, 3.0)):
        stars.append(_make_star(ds, target_resolution_lnP=0.3, core_temperature=Tc, Q_heating_shallow=Q,
                                number_epochs=3, basic_epoch_Mdots=[2.0e17, 0.0, 0.0],
                                basic_epoch_boundaries=[-400.0, 0.0, 100.0, 2000.0],
                                fix_atmosphere_temperature_when_accreting=False))
    ds.create_models(stars)
    lnT0 = [s.array("lnT").copy() for s in stars]
    ds.evolve_models(stars)
    for s, y0, Q in zip(stars, lnT0, (0.0, 1.7, 3.0)):
        ref = oracle_evolve_star(oracle, s, y0, heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, Q],
                                 ctrl=dict(turn_on_extra_heating=1, fix_atmosphere_temperature_when_accreting=0))
        t, Teff, Qb = s.monitors()
        assert rel(Teff, ref["Teff_monitor"]) < 1e-6, (Teff, ref["Teff_monitor"])
        h = s.history()
        assert len(h["model"]) == s.counters()[3] + 3  # one row per accepted step + the first point of each epoch
        s.free()


def test_evolve_warp_mapping_matches_cta_mapping(nscool, oracle):
    """The warp-per-star form of the evolve kernel (zones in registers, partitioned solve with cyclic reduction over
    warp shuffles) runs the same RODASP driver on the same right-hand side as the CTA-per-star form; only the
    elimination order of the tridiagonal solve differs, so light curves agree far below the integration tolerance and
    the solver statistics are identical.  Ragged batch: 2, 4 and 8 zones per lane in one call."""
    ds = nscool
    from dstar_b200._capi import lib
    out = {}
    for mapping in (1, 2):
        assert lib().dstar_NScool_set_evolve_mapping(mapping) == 0
        try:
            stars = [_make_star(ds, target_resolution_lnP=r, core_temperature=Tc, number_epochs=3,
                                basic_epoch_Mdots=[2.0e17, 0.0, 0.0], basic_epoch_boundaries=[-400.0, 0.0, 100.0, 2000.0])
                     for r, Tc in ((0.6, 5e7), (0.3, 9.2e7), (0.12, 2e8))]
            ds.create_models(stars)
            lnT0 = [s.array("lnT").copy() for s in stars]
            ds.evolve_models(stars)
            out[mapping] = [(s.monitors(), s.counters().copy(), s.array("lnT").copy(), s.get_int("nz")) for s in stars]
            if mapping == 2:
                ref = oracle_evolve_star(oracle, stars[2], lnT0[2], heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, 0.1],
                                         T_atm=2.4e8, ctrl=dict(turn_on_extra_heating=1))
                assert rel(stars[2].monitors()[1], ref["Teff_monitor"]) < 1e-6
                np.testing.assert_array_equal(stars[2].counters(), ref["counters"])
            for s in stars:
                s.free()
        finally:
            lib().dstar_NScool_set_evolve_mapping(0)
    for (m1, c1, y1, nz), (m2, c2, y2, _) in zip(out[1], out[2]):
        assert rel(m2[1], m1[1]) < 1e-9, (nz, m1[1], m2[1])
        assert np.max(np.abs(y1 - y2)) < 1e-9
        np.testing.assert_array_equal(c1, c2)
    print("zones per star:", [o[3] for o in out[1]])


def test_crust_composition_lookup_parity(ctx, oracle):
    """north_star (b): composition lookups on a device-resident table vs dStar_crust_get_composition_info."""
    import dstar_b200 as ds
    rng = np.random.default_rng(SEED + 3)
    nv = 400
    lgP_tab = np.linspace(22.0, 33.5, nv)
    Ziso = np.concatenate([[0.0], NET_Z]); Aiso = np.concatenate([[1.0], NET_A])
    Y = np.zeros((len(Ziso), nv))
    for j in range(nv):   # smooth random composition drifting with depth
        w = np.abs(np.sin(0.01 * j * np.arange(1, len(Ziso) + 1) + rng.uniform(0, 0.01)))
        w[0] *= 3.0 * j / nv
        Y[:, j] = w / np.sum(w * Aiso)
    Ycoef = np.concatenate([ds.host_interp_pm(lgP_tab, Y[i]).ravel() for i in range(len(Ziso))])
    lgP = rng.uniform(21.5, 34.0, 3000)
    ionic_g, Yion_g, nch = ctx.crust_composition_info(lgP_tab, Ycoef, Ziso, Aiso, lgP)
    import ctypes as C
    from oracle_lib import _ptr, f64, ip
    ionic_o = np.zeros((len(lgP), 11)); Yion_o = np.zeros((len(lgP), nch)); n2 = C.c_int32()
    a = [f64(x) for x in (lgP_tab, Ycoef, Ziso, Aiso, lgP)]
    oracle.lib.dso_crust_composition(nv, len(Ziso), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), len(lgP),
                                     _ptr(a[4]), _ptr(ionic_o), _ptr(Yion_o), C.byref(n2))
    assert n2.value == nch == len(NET_Z)
    assert rel(ionic_g, ionic_o) < 1e-12
    assert np.max(np.abs(Yion_g - Yion_o)) < 1e-15


def test_call_order_errors(ctx):
    """ierr conventions: state errors are negative return codes, never silent fallbacks."""
    import ctypes as C

    import dstar_b200 as ds
    b = ds.Batch(ctx, [0, 8], 2)
    with pytest.raises(ds.DStarError):
        b.micro_tables()          # structure not uploaded
    with pytest.raises(ds.DStarError):
        b.evolve()
    b.close()
    with pytest.raises(ds.DStarError):
        ds.Batch(ctx, [0, 2], 2)  # fewer than 3 zones


@pytest.mark.gpu
def test_micro_tables_redo_pass_matches(nscool):
    """The coefficient pass evaluates the cell EOS with branch-free divisions and re-evaluates flagged zones
    with plain IEEE divisions in a second launch; forcing every zone through that second pass must reproduce
    the tables bit for bit (the two division forms are the same arithmetic)."""
    s = _make_star(nscool)
    s.create_model()
    ref = {k: s.array(k).copy() for k in ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK")}
    s.free()
    from dstar_b200._capi import lib
    lib().dstar_debug_force_redo(1)
    try:
        s2 = _make_star(nscool)
        s2.create_model()
        for k, v in ref.items():
            assert np.array_equal(s2.array(k).view(np.uint64), v.view(np.uint64)), k
        s2.free()
    finally:
        lib().dstar_debug_force_redo(0)
