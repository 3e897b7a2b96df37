"""Seam B below the light-curve level: the device right-hand side and Jacobian rows against the oracle's
get_derivatives / get_jacobian (NScool/private/NScool_evolve.f:237-359) at given ln T, and the per-step history the
kernel keeps (evaluate_timestep, :138-235: model, t, dt, T_eff, L_surf, L_nu, L_nuc and the terminal line's extrema)
against the oracle's solout trace, row by row."""
import numpy as np
import pytest

from oracle_lib import RhsEvaluator
from parity_helpers import evolve_in_kwargs, oracle_evolve_star

pytestmark = pytest.mark.gpu

HEAT = [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, 0.1]
LN10 = np.log(10.0)


@pytest.fixture(scope="module")
def nscool():
    import dstar_b200 as ds
    ds.NScool_init()
    yield ds
    ds.NScool_shutdown()


def make_star(ds, **kw):
    base = dict(target_resolution_lnP=0.2, number_epochs=2, basic_epoch_Mdots=[1.0e17, 0.0],
                basic_epoch_boundaries=[0.0, 50.0, 1050.0], core_temperature=1.4e8,
                atmosphere_temperature_when_accreting=2.4e8, turn_on_extra_heating=True, Q_heating_shallow=0.1,
                lgP_min_heating_shallow=27.5, lgP_max_heating_shallow=28.5, which_neutron_1S0_gap="sfb03",
                fix_Qimp=True, Qimp=1.0)
    base.update(kw)
    return ds.NScool(None, **base)


def batch_from_star(ds, ctx, s, lnT, heat=HEAT, T_atm=2.4e8):
    """A kernel-level batch holding one product-built star (structure, its own tables, evolve inputs) at ln T."""
    nz = s.get_int("nz"); nch = s.get_int("ncharged")
    b = ds.Batch(ctx, np.array([0, nz]), nch)
    b.upload_structure(*[s.array(k) for k in ("rho", "rho_bar", "P", "P_bar", "dm", "ionic", "ionic_bar", "Yion", "Yion_bar",
                                              "Zion", "Aion")])
    b.upload_tables(*[s.array(k) for k in ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK")])
    eb = s.array("epoch_boundaries").reshape(-1, 1); em = s.array("epoch_Mdots").reshape(-1, 1)
    b.upload_evolve(*[s.array(k) for k in ("dm_bar", "area", "ePhi", "e2Phi", "ePhi_bar", "e2Phi_bar")],
                    s.array("atm_lgTb"), s.array("atm_lgTeff"), s.array("atm_lgflux"), [0], np.array(heat), [T_atm], lnT, eb, em)
    return b


@pytest.mark.parametrize("insulating", [False, True])
def test_device_rhs_and_jacobian_match_oracle(nscool, oracle, insulating):
    """f(y) and the three Jacobian diagonals of every zone, accreting (fixed atmosphere) and cooling epochs, at the
    initial isothermal profile and at the evolved (non-isothermal) one: relative deviation <= 1e-12 of the row scale.
    f itself is the small difference of neighbouring fluxes (they cancel exactly in an isothermal crust), and the device
    forms it with precomputed reciprocals of the zone constants, so its deviation is measured against the
    scale of the terms it is the difference of -- the row scale of the Jacobian -- not against f."""
    ds = nscool
    s = make_star(ds)
    s.create_model()
    y_iso = s.array("lnT").copy()
    s.evolve_model()
    y_evolved = s.array("lnT").copy()
    ctrl = ds.EvolveCtrl.defaults()
    ctrl.turn_on_extra_heating = 1
    octrl = dict(turn_on_extra_heating=1)
    if insulating:
        ctrl.fix_core_temperature = 0; ctrl.make_inner_boundary_insulating = 1
        octrl.update(fix_core_temperature=0, make_inner_boundary_insulating=1)
    ctx = ds.Context(0, gaps=("ns", "sfb03", "t72"))
    worst = 0.0
    try:
        for y in (y_iso, y_evolved):
            b = batch_from_star(ds, ctx, s, y)
            for epoch in (0, 1):
                g = b.eval_rhs(epoch, ctrl)
                ev = RhsEvaluator(oracle, evolve_in_kwargs(s, epoch, ctrl=octrl, heat=HEAT, T_atm=2.4e8))
                o = ev(y)
                scale = np.maximum.reduce([np.abs(o["di"]), np.abs(o["lo"]), np.abs(o["up"])]) + 1e-300
                fscale = np.max(np.abs(o["f"])) + 1e-300
                fdev = np.max(np.abs(g["f"] - o["f"]) / (scale + fscale))
                assert fdev < 1e-12, (epoch, fdev)
                for nm in ("lo", "di", "up"):
                    dev = np.max(np.abs(g[nm] - o[nm]) / scale)
                    worst = max(worst, dev)
                    assert dev < 1e-12, (nm, epoch, dev)
                assert abs(g["Teff"][0] / o["Teff"] - 1.0) < 1e-13 and abs(g["Lsurf"][0] / o["Lsurf"] - 1.0) < 1e-13
                assert abs(g["dlnLsdlnT"][0] - o["dlnLsdlnT"]) < 1e-12
            b.close()
    finally:
        ctx.close()
        s.free()
    print("worst Jacobian deviation / row scale:", worst)


@pytest.mark.parametrize("mapping", [1, 2])
def test_history_rows_match_oracle_trace(nscool, oracle, mapping, tmp_path):
    """Every history row the kernel keeps (both thread mappings) against the oracle's solout trace: model numbers and
    times exactly, dt/T_eff/L_surf/L_nu/L_nuc to 1e-6, and the terminal line's extrema of ln T with their pressures."""
    ds = nscool
    from dstar_b200._capi import lib
    lib().dstar_NScool_set_evolve_mapping(mapping)
    try:
        s = make_star(ds, number_epochs=3, basic_epoch_Mdots=[1.0e17, 0.0, 3.0e16], basic_epoch_boundaries=[0.0, 50.0, 1050.0, 1100.0])
        s.create_model()
        y0 = s.array("lnT").copy()
        s.evolve_model()
        h = s.history(); term = s.terminal()
        P = s.array("P")
        ref = oracle_evolve_star(oracle, s, y0, heat=HEAT, T_atm=2.4e8, ctrl=dict(turn_on_extra_heating=1), trace_cap=4096)
        rows = {k: np.concatenate([t[k] for t in ref["trace"]]) for k in ("model", "tsec", "dt", "Teff", "Lsurf", "Lnu", "Lnuc")}
        lnT_rows = np.concatenate([t["lnT_rows"] for t in ref["trace"]])
        n = len(rows["model"])
        assert len(h["model"]) == n and n == s.counters()[3] + 3      # accepted steps + the first point of each epoch
        np.testing.assert_array_equal(h["model"], rows["model"])
        assert np.max(np.abs(h["time"] * 86400.0 - rows["tsec"]) / np.maximum(rows["tsec"], 1.0)) < 1e-9
        assert np.max(np.abs(term["dt"] - rows["dt"]) / np.maximum(rows["dt"], 1e-30)) < 1e-6
        for nm, col in (("lg_Teff", "Teff"), ("lg_Lsurf", "Lsurf"), ("lg_Lnu", "Lnu")):
            assert np.max(np.abs(h[nm] - np.log10(rows[col]))) < 1e-6 / LN10 * 2, nm
        acc = rows["Lnuc"] > 0
        assert np.max(np.abs(h["lg_Lnuc"][acc] - np.log10(rows["Lnuc"][acc]))) < 1e-9 and np.all(h["lg_Lnuc"][~acc] == -99.0)
        # terminal extrema (NScool_terminal.f:53-56)
        assert np.max(np.abs(term["lgTmax"] - lnT_rows.max(axis=1) / LN10)) < 1e-6
        assert np.max(np.abs(term["lgTmin"] - lnT_rows.min(axis=1) / LN10)) < 1e-6
        np.testing.assert_allclose(term["lgP_Tmax"], np.log10(P[lnT_rows.argmax(axis=1)]), rtol=0, atol=1e-12)
        np.testing.assert_allclose(term["lgP_Tmin"], np.log10(P[lnT_rows.argmin(axis=1)]), rtol=0, atol=1e-12)
        # the files: history.data as NScool_history.f writes it, the terminal log as NScool_terminal.f does
        assert s.write_history(str(tmp_path / "history.data")) == 0
        assert s.write_terminal(str(tmp_path / "terminal.txt")) == 0
        lines = open(tmp_path / "terminal.txt").read().split("\n")
        body = [l for l in lines if l.strip() and not l.strip().startswith("model")]
        assert len(body) == n and len(body[0]) == 11 * 15
        assert sum(1 for l in lines if l.strip().startswith("model")) >= 3     # a header at the start of each epoch
        s.free()
    finally:
        lib().dstar_NScool_set_evolve_mapping(0)


def test_history_truncation_is_reported(nscool, tmp_path):
    ds = nscool
    s = make_star(ds, trace_capacity=5)
    s.create_model(); s.evolve_model()
    assert len(s.history()["model"]) == 5
    assert s.write_history(str(tmp_path / "h.data")) == 1          # written, but flagged as truncated
    s.free()
    s = make_star(ds, trace_capacity=0)
    s.create_model(); s.evolve_model()
    with pytest.raises(ds.DStarError):
        s.write_history(str(tmp_path / "h2.data"))
    s.free()


def test_mixed_controls_in_one_batch_are_refused(nscool):
    """Settings the kernels take once per batch (EOS/conductivity/neutrino switches, gap tables and scales, integration
    controls) must agree across the stars of a batch -- as the nuclide network must."""
    ds = nscool
    a = make_star(ds); b = make_star(ds, which_neutron_1S0_gap="gc")
    with pytest.raises(ds.DStarError):
        ds.create_models([a, b])
    b.free()
    b = make_star(ds, use_neutron_conductivity=True)
    with pytest.raises(ds.DStarError):
        ds.create_models([a, b])
    b.free()
    b = make_star(ds, integration_tolerance=1.0e-5)
    ds.create_models([a, b])
    with pytest.raises(ds.DStarError):
        ds.evolve_models([a, b])
    a.free(); b.free()
