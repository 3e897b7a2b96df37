"""bc09 envelope tables (the reference's default atm_model) generated on the device, against the CPU oracle's
knot-independent form (oracle/atm.cpp, Bc09Settings::tight -- same algorithm, same operation order): every knot of
lgTb(lgTeff) must agree to 1e-10 (observed: identical bits), the photospheres too."""
import numpy as np
import pytest

from parity_helpers import oracle_evolve_star, oracle_micro_tables

pytestmark = pytest.mark.gpu

G = 2.43e14
PB = 4.3e13 * G


@pytest.fixture(scope="module")
def ctx():
    import dstar_b200 as ds
    c = ds.Context(0, gaps=("ns", "gc", "t72"))
    yield c
    c.close()


def test_bc09_knots_match_oracle(ctx, oracle):
    # test_atm.f's two light-element columns, plus a case in which trial stages of dop853 land on temperatures whose
    # radiation pressure exceeds P (no density solves the EOS: the "retry with smaller timestep" path); all 256 knots
    lgTeff = 5.5 + np.arange(256) * 1.5 / 255.0
    grav = np.array([G, G, 1.92e14]); Plight = np.array([2.43e19, 2.43e23, 1.92e14 * 1e4]); Pb = np.array([PB, PB, 1.0e27])
    _, lgTb, lgflux, ierr, info = ctx.atm_bc09_Teff(grav, Plight, Pb, lgTeff, with_info=True)
    assert np.all(ierr == 0)
    for t in range(3):
        rc, ref_Tb, ref_fl, ref_info = oracle.bc09_Teff(grav[t], Plight[t], Pb[t], lgTeff, tight=True)
        assert rc == 0
        np.testing.assert_allclose(info[t, :, :3], ref_info[:, :3], rtol=1e-10)   # rho_ph, P_ph, kappa_ph
        np.testing.assert_array_equal(info[t, :, 3], ref_info[:, 5])              # same number of EOS evaluations
        np.testing.assert_allclose(lgTb[t], ref_Tb, rtol=0, atol=1e-10)
        np.testing.assert_allclose(lgflux[t], ref_fl, rtol=0, atol=1e-12)
        print("bc09 table %d: max |lgTb - oracle| = %.2e" % (t, np.abs(lgTb[t] - ref_Tb).max()))


def test_bc09_golden_through_device(ctx, oracle):
    """The reference's own golden rows (dStar_atm/test/sample_output.data, bc09 blocks), table from the DEVICE."""
    from test_oracle_atm import _bc09_blocks
    import dstar_b200 as ds
    gold = _bc09_blocks()
    lgTeff, lgTb, lgflux, ierr = ctx.atm_bc09_Teff([G, G], [2.43e19, 2.43e23], [PB, PB])
    assert np.all(ierr == 0)
    for t, g in enumerate(gold):
        f = ds.host_interp_pm(lgTb[t], lgTeff)
        v, s = oracle.interp_value_and_slope(lgTb[t], f, np.clip(g[:, 0], 7.0, 9.5))
        np.testing.assert_allclose(v, g[:, 1], atol=1.0e-5)
        rel = np.abs(s / g[:, 2] - 1.0)
        assert np.sum(rel > 1.0e-3) <= 1 and rel.max() < 5.0e-3, rel


def test_nscool_default_atmosphere_is_bc09(oracle):
    """create_model with the default atm_model: the star's atmosphere table equals do_generate_atm_table('bc09') of
    the oracle for the star's own (grav, Plight, Psurf); a second star with the same triple reuses the table; the
    light curve through that boundary matches the oracle's integration."""
    import dstar_b200 as ds
    ds.NScool_init()
    try:
        kw = dict(target_resolution_lnP=0.2, number_epochs=2, basic_epoch_Mdots=[1.0e17, 0.0],
                  basic_epoch_boundaries=[0.0, 50.0, 1050.0], core_temperature=1.4e8,
                  atmosphere_temperature_when_accreting=2.4e8, fix_atmosphere_temperature_when_accreting=False,
                  which_neutron_1S0_gap="sfb03", fix_Qimp=True, Qimp=1.0, lg_atm_light_element_column=8.0)
        a = ds.NScool(None, **kw)   # atm_model not set: the default is bc09 (NScool/defaults/controls.defaults:149)
        b = ds.NScool(None, **dict(kw, core_temperature=9.0e7))
        ds.create_models([a, b])
        lnT0 = a.array("lnT").copy()
        g, Ps = a.get_real("grav"), a.get_real("Psurf")
        rc, lgTb, lgTeff4, lgflux4 = oracle.bc09_table(g, g * 10.0 ** 8.0, Ps, tight=True)
        assert rc == 0
        for got, ref in ((a.array("atm_lgTb"), lgTb), (a.array("atm_lgTeff"), lgTeff4.ravel()),
                         (a.array("atm_lgflux"), lgflux4.ravel())):
            scale = np.maximum(np.abs(ref), 1.0)
            assert np.max(np.abs(got - ref) / scale) < 1e-9, np.max(np.abs(got - ref) / scale)
        np.testing.assert_array_equal(a.array("atm_lgTb"), b.array("atm_lgTb"))
        ds.evolve_models([a, b])
        t, Teff, Qb = a.monitors()
        ref = oracle_evolve_star(oracle, a, lnT0, T_atm=2.4e8, ctrl=dict(fix_atmosphere_temperature_when_accreting=0))
        np.testing.assert_allclose(Teff, ref["Teff_monitor"], rtol=1e-6)
    finally:
        ds.NScool_shutdown()
