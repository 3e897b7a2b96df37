"""The synthetic workloads of BASELINE.json's configs (dstar_b200/workloads.py) at test sizes: parity against the
CPU oracle, and size-independent properties (batch invariance, idempotence, ragged batches, extreme zone counts)."""
import os
import sys

import numpy as np
import pytest

from parity_helpers import oracle_evolve_star, oracle_micro_tables

pytestmark = pytest.mark.gpu

TABS = ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK")


@pytest.fixture(scope="module")
def nscool():
    import dstar_b200 as ds
    ds.NScool_init()
    yield ds
    ds.NScool_shutdown()


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def check_tables(s, ref, tol=1e-10):
    for name in TABS:
        g = s.array(name).reshape(-1, 4); r = ref[name].reshape(-1, 4)
        assert np.max(np.abs(g[:, 0] - r[:, 0])) < tol, name


def test_hooks_sweep_parity(nscool, oracle):
    """configs[1] (custom_Tc_Qimp): 2 x 2 grid with the example's alt_Qimp / alt_sf hooks -- tables (Tc and Q as
    the hooks set them per zone) and light curves against the oracle."""
    from dstar_b200 import workloads
    stars = workloads.custom_Tc_Qimp(4, resolution=0.4)
    nscool.create_models(stars)
    y0 = [s.array("lnT").copy() for s in stars]
    nscool.evolve_models(stars)
    for s, lnT0 in zip(stars, y0):
        ref = oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72"), Tc_cell=s.array("Tc_cell"),
                                  Tc_face=s.array("Tc_face"))
        check_tables(s, ref)
        ev = oracle_evolve_star(oracle, s, lnT0, heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 1.7],
                                ctrl=dict(turn_on_extra_heating=1, fix_atmosphere_temperature_when_accreting=0))
        t, Teff, Qb = s.monitors()
        assert len(Teff) == 9 and rel(Teff, ev["Teff_monitor"]) < 1e-6
    # the hooks matter: pasta Qimp (second control) changes the face tables, not the cell tables
    a, b = stars[0], stars[1]           # same Tc, different pasta Qimp
    assert np.array_equal(bits(a.array("tab_lnCp")), bits(b.array("tab_lnCp")))
    assert not np.array_equal(bits(a.array("tab_lnK")), bits(b.array("tab_lnK")))
    for s in stars:
        s.free()


def test_fit_lightcurve_walkers(nscool, oracle):
    """configs[2]: MCMC walkers over (Mdot, Q_heating_shallow, Qimp), 7 monitor epochs; two of them against the oracle."""
    from dstar_b200 import workloads
    stars = workloads.fit_lightcurve(8, resolution=0.3)
    nscool.create_models(stars)
    y0 = [s.array("lnT").copy() for s in stars]
    nscool.evolve_models(stars)
    rng = np.random.default_rng(workloads.SEED + 2)
    rng.uniform(0, 1, 8)
    Qh = rng.uniform(0.0, 3.0, 8)
    curves = []
    for k, s in enumerate(stars):
        t, Teff, Qb = s.monitors()
        assert len(Teff) == 7 and np.all(np.isfinite(Teff)) and np.all(Teff > 0)
        curves.append(Teff)
        if k in (0, 5):
            check_tables(s, oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72")))
            ev = oracle_evolve_star(oracle, s, y0[k], heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, float(Qh[k])],
                                    ctrl=dict(turn_on_extra_heating=1, fix_atmosphere_temperature_when_accreting=0))
            assert rel(Teff, ev["Teff_monitor"]) < 1e-6
    assert len({tuple(np.round(c, 3)) for c in curves}) == 8      # eight different walkers, eight different curves
    for s in stars:
        s.free()


def test_accretion_history_neutron_channels(nscool, oracle):
    """configs[3]: two accretion cycles (10 epochs) with neutron + superfluid-phonon conductivity and shell Urca."""
    from dstar_b200 import workloads
    stars = workloads.accretion(2, resolution=0.6)
    nscool.create_models(stars)
    y0 = [s.array("lnT").copy() for s in stars]
    nscool.evolve_models(stars)
    s = stars[0]
    ref = oracle_micro_tables(oracle, s, gaps=("ns", "gc", "t72"), cond_ctrl=(1, 1, 1, 1, 10.0, 9.0),
                              shell_Urca=(1, 1.0, 28.5))
    check_tables(s, ref)
    ev = oracle_evolve_star(oracle, s, y0[0], heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 1.0], T_atm=5.0e8,
                            ctrl=dict(turn_on_extra_heating=1))
    t, Teff, Qb = s.monitors()
    assert len(Teff) == 10 and rel(Teff, ev["Teff_monitor"]) < 1e-6
    for s in stars:
        s.free()


def _star(ds, res, **kw):
    base = dict(target_resolution_lnP=res, number_epochs=2, basic_epoch_Mdots=[1.0e17, 0.0],
                basic_epoch_boundaries=[0.0, 50.0, 1050.0], core_temperature=1.4e8,
                atmosphere_temperature_when_accreting=2.4e8, turn_on_extra_heating=True, Q_heating_shallow=0.1,
                lgP_min_heating_shallow=27.5, lgP_max_heating_shallow=28.5, which_neutron_1S0_gap="sfb03",
                fix_Qimp=True, Qimp=1.0)
    base.update(kw)
    return ds.NScool(None, **base)


def test_ragged_batch_is_bitwise_batch_invariant(nscool):
    """Stars with different zone counts in ONE batch: every table entry, monitor and solver counter equals the
    star's own single-star run bit for bit (no cross-star coupling, no dependence on the launch geometry)."""
    ds = nscool
    res = (0.6, 0.17, 0.31, 0.17)
    batch = [_star(ds, r, core_temperature=1.0e8 * (1 + i)) for i, r in enumerate(res)]
    ds.create_models(batch)
    ds.evolve_models(batch)
    assert len({s.get_int("nz") for s in batch}) == 3
    for i, r in enumerate(res):
        solo = _star(ds, r, core_temperature=1.0e8 * (1 + i))
        solo.create_model()
        solo.evolve_model()
        for name in TABS + ("lnT",):
            assert np.array_equal(bits(batch[i].array(name)), bits(solo.array(name))), (i, name)
        for a, b in zip(batch[i].monitors(), solo.monitors()):
            assert np.array_equal(bits(a), bits(b))
        np.testing.assert_array_equal(batch[i].counters(), solo.counters())
        solo.free()
    for s in batch:
        s.free()


@pytest.mark.parametrize("res", [0.0252, 3.0])
def test_extreme_zone_counts(nscool, oracle, res):
    """Largest grid the kernels accept (nz ~ 500 of 512: the 512-thread evolve kernel) and a crust of a handful of zones."""
    s = _star(nscool, res)
    s.create_model()
    nz = s.get_int("nz")
    assert (nz > 480 and nz <= 512) if res < 1 else (3 <= nz <= 6)
    lnT0 = s.array("lnT").copy()
    s.evolve_model()
    check_tables(s, oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72")))
    ev = oracle_evolve_star(oracle, s, lnT0, heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, 0.1], T_atm=2.4e8,
                            ctrl=dict(turn_on_extra_heating=1))
    assert rel(s.monitors()[1], ev["Teff_monitor"]) < 1e-6
    s.free()


def test_too_many_zones_is_refused(nscool):
    import dstar_b200 as ds
    s = _star(nscool, 0.02)      # ~633 zones > 512
    with pytest.raises(ds.DStarError):
        s.create_model()
    s.free()


def test_full_size_sweep_properties(nscool, oracle):
    """configs[1] at BASELINE size (256 stars x 253 zones): everything converges, the coefficient pass is idempotent,
    cell tables depend on the core Tc only through ... nothing (same crust), face tables only on pasta Qimp, a
    star pulled out of the batch reproduces its in-batch result bit for bit, and six stars spread over the grid equal
    the oracle at full resolution (every table entry to 1e-10, T_eff at the nine epochs to 1e-6, same solver counters)."""
    from dstar_b200 import workloads
    ds = nscool
    stars = workloads.custom_Tc_Qimp(256, resolution=0.05)
    ds.create_models(stars)
    first = {n: stars[37].array(n).copy() for n in TABS}
    ds.create_models(stars)                                    # idempotence of the coefficient pass
    for n in TABS:
        assert np.array_equal(bits(stars[37].array(n)), bits(first[n])), n
    picks = (0, 15, 85, 119, 170, 255)
    y0 = {i: stars[i].array("lnT").copy() for i in picks}
    ds.evolve_models(stars)
    for i in picks:
        s = stars[i]
        ref = oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72"), Tc_cell=s.array("Tc_cell"), Tc_face=s.array("Tc_face"))
        check_tables(s, ref)
        ev = oracle_evolve_star(oracle, s, y0[i], heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 1.7],
                                ctrl=dict(turn_on_extra_heating=1, fix_atmosphere_temperature_when_accreting=0))
        assert rel(s.monitors()[1], ev["Teff_monitor"]) < 1e-6, i
        np.testing.assert_array_equal(s.counters(), ev["counters"])
    Teff = np.array([s.monitors()[1] for s in stars])
    assert Teff.shape == (256, 9) and np.all(np.isfinite(Teff)) and np.all(Teff > 1e5) and np.all(Teff < 1e7)
    nz = stars[0].get_int("nz")
    assert 250 <= nz <= 256
    # grid structure: star index = 16 * i_Tc + j_Q
    for j in (1, 9):
        assert np.array_equal(bits(stars[j].array("tab_lnCp")), bits(stars[0].array("tab_lnCp")))
        assert np.array_equal(bits(stars[16 * 3 + j].array("tab_lnK")), bits(stars[j].array("tab_lnK")))
    # hotter core => hotter surface at the last (longest-cooled) epoch, for every Qimp column
    last = Teff[:, -1].reshape(16, 16)
    assert np.all(np.diff(last, axis=0) > 0)
    # batch invariance at full size
    solo = workloads.custom_Tc_Qimp(256, resolution=0.05)[200]
    solo.create_model(); solo.evolve_model()
    for a, b in zip(stars[200].monitors(), solo.monitors()):
        assert np.array_equal(bits(a), bits(b))
    np.testing.assert_array_equal(stars[200].counters(), solo.counters())
    for s in stars:
        s.free()


def test_shared_tables_are_bit_identical(nscool):
    """share_identical_tables: configs[1] as a 4 x 4 (core Tc, pasta Qimp) grid needs ONE cell-table set (the crust is
    the same) and four face-table sets (one per Qimp); every star still reports its own tables, bit for bit those of
    the unshared pass, and evolves to the same bits.  A batch of walkers with different Qimp shares cells only."""
    from dstar_b200 import workloads
    ds = nscool
    plain = workloads.custom_Tc_Qimp(16, resolution=0.3)
    ds.create_models(plain)
    assert ds.table_sets() == (16, 16)
    ds.evolve_models(plain)
    shared = workloads.custom_Tc_Qimp(16, resolution=0.3)
    for s in shared:
        s.set_control("share_identical_tables", True)
    ds.create_models(shared)
    assert ds.table_sets() == (1, 4)
    ds.evolve_models(shared)
    for a, b in zip(plain, shared):
        for n in TABS:
            assert np.array_equal(bits(a.array(n)), bits(b.array(n))), n
        assert a.array("rho_bar")[0] == b.array("rho_bar")[0]
        for x, y in zip(a.monitors(), b.monitors()):
            assert np.array_equal(bits(x), bits(y))
        np.testing.assert_array_equal(a.counters(), b.counters())
        assert np.array_equal(bits(a.array("lnT")), bits(b.array("lnT")))
    for s in plain + shared:
        s.free()
    walkers = workloads.fit_lightcurve(6, resolution=0.4)
    for s in walkers:
        s.set_control("share_identical_tables", True)
    ds.create_models(walkers)
    assert ds.table_sets() == (1, 6)
    # mixed zone counts never share
    ragged = [_star(ds, r, share_identical_tables=True) for r in (0.4, 0.4, 0.25)]
    ds.create_models(ragged)
    assert ds.table_sets() == (2, 2)
    ds.evolve_models(ragged)
    assert np.array_equal(bits(ragged[0].monitors()[1]), bits(ragged[1].monitors()[1]))
    for s in walkers + ragged:
        s.free()


def test_abuntime_composition_many_species(nscool, oracle):
    """configs[4] flavour (INT-16-2b-demo): crust composition from an abuntime-format table -- ~60 nuclides above the
    abundance threshold per zone (the ion-mixture loop, ion.f:72-83, dominates) blended into HZ90 at depth."""
    from dstar_b200 import workloads
    stars = workloads.int16_demo(2, resolution=0.5)
    nscool.create_models(stars)
    s = stars[0]
    nch = s.get_int("ncharged")
    assert nch >= 60
    Yion = s.array("Yion").reshape(-1, nch)
    assert (Yion[0] > 1e-9).sum() >= 40 and (Yion[-1] > 1e-9).sum() <= 2      # many species outside, HZ90 at the bottom
    y0 = [x.array("lnT").copy() for x in stars]
    nscool.evolve_models(stars)
    check_tables(s, oracle_micro_tables(oracle, s, gaps=("ns", "sfb03", "t72")))
    rng = np.random.default_rng(workloads.SEED + 4)
    rng.uniform(0, 1, 2)
    Qh = rng.uniform(0.0, 3.0, 2)
    ev = oracle_evolve_star(oracle, s, y0[0], heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, float(Qh[0])],
                            ctrl=dict(turn_on_extra_heating=1, fix_atmosphere_temperature_when_accreting=0))
    t, Teff, Qb = s.monitors()
    assert len(Teff) == 10 and rel(Teff, ev["Teff_monitor"]) < 1e-6
    for x in stars:
        x.free()


def test_second_outburst_recovers_from_oversized_first_step(nscool, oracle):
    """examples/accretion at full resolution: the second accretion cycle starts with the step size left over from
    3e4 days of cooling; the first trial steps overflow (NaN error norm) and must simply be halved until one fits,
    on the device as in the oracle."""
    from dstar_b200 import workloads
    stars = workloads.accretion(2, resolution=0.05)
    for s in stars:   # the electron-only conductivity keeps the oracle's share of this test short
        s.set_control("use_neutron_conductivity", False); s.set_control("use_superfluid_phonon_conductivity", False)
    nscool.create_models(stars)
    y0 = [s.array("lnT").copy() for s in stars]
    nscool.evolve_models(stars)
    for k, s in enumerate(stars):
        t, Teff, Qb = s.monitors()
        assert len(Teff) == 10 and np.all(Teff > 0)
        ev = oracle_evolve_star(oracle, s, y0[k], heat=[26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 1.0], T_atm=5.0e8,
                                ctrl=dict(turn_on_extra_heating=1))
        assert rel(Teff, ev["Teff_monitor"]) < 1e-6
        np.testing.assert_array_equal(s.counters(), ev["counters"])
        s.free()


def test_profiles_from_device_capture(nscool, tmp_path):
    """do_write_profile (NScool_profile.f): ln T of every zone captured on the device at mod(model, interval) == 0;
    the other columns rebuilt from the tables.  Checked: model numbering, the last profile against the final state,
    columns against a direct numpy evaluation of the same splines, file layout (header, 20 columns, manifest)."""
    s = _star(nscool, 0.3, write_interval_for_profile=3, suppress_first_step_output=False)
    s.create_model()
    s.evolve_model()
    kept, made = s.profile_count()
    naccpt = s.counters()[3]
    assert kept == made and kept >= (naccpt // 3) and kept <= naccpt // 3 + 4
    nz = s.get_int("nz")
    profs = [s.profile(i) for i in range(kept)]
    models = [q["model"] for q in profs]
    assert all(m % 3 == 0 for m in models) and models == sorted(models) and len(set(models)) == len(models)
    times = [q["time"] for q in profs]
    assert times == sorted(times) and times[-1] <= 1050.0 + 1e-9
    q = profs[-1]
    # columns: structure straight from the star, microphysics from the tables at the profile's temperature
    assert np.array_equal(q["zone"], np.arange(1, nz + 1))
    assert np.array_equal(q["dm"], s.array("dm")) and np.array_equal(q["density"], s.array("rho"))
    assert np.array_equal(q["pressure"], s.array("P_bar"))
    tab = s.array("tab_lnT"); lnT = np.log(q["temperature"])

    def spline(name, x):
        f = s.array(name).reshape(nz, 128, 4)
        k = np.clip(np.searchsorted(tab, x, side="right") - 1, 0, 126)
        dx = x - tab[k]
        c = f[np.arange(nz), k]
        return np.exp(c[:, 0] + dx * (c[:, 1] + dx * (c[:, 2] + dx * c[:, 3])))
    assert rel(q["cp"], spline("tab_lnCp", lnT)) < 1e-12
    assert rel(q["Gamma"], spline("tab_lnGamma", lnT)) < 1e-12
    assert rel(q["eps_nu"], spline("tab_lnEnu", lnT)) < 1e-12
    assert q["cooling_lum"] == pytest.approx(float(np.dot(q["eps_nu"], q["dm"])), rel=1e-12)
    assert q["accretion_rate"] == 0.0 and np.all(q["eps_nuc"] == 0.0) and q["heating_lum"] == 0.0   # quiescent epoch
    assert profs[0]["accretion_rate"] == 1.0e17 and profs[0]["heating_lum"] > 0.0                  # outburst epoch
    T = q["temperature"]; dm = s.array("dm"); dmb = s.array("dm_bar")
    Tbar = np.concatenate([[T[0]], 0.5 * (T[:-1] * dm[1:] + T[1:] * dm[:-1]) / dmb[1:]])
    assert rel(q["K"], spline("tab_lnK", np.log(Tbar))) < 1e-12
    ePhi = s.array("ePhi"); area = s.array("area")
    Lref = -(area[1:] ** 2) * s.array("rho_bar")[1:] * q["K"][1:] * (ePhi[:-1] * T[:-1] - ePhi[1:] * T[1:]) \
        / s.array("ePhi_bar")[1:] / dmb[1:]
    assert rel(q["luminosity"][1:], Lref) < 1e-12 and q["luminosity"][0] > 0      # heat flows out at the surface
    # if the last accepted model is a profile model it equals the final state
    if models[-1] == s.get_int("model"):
        assert np.array_equal(lnT, s.array("lnT"))
    # files
    s.write_profiles(str(tmp_path))
    man = (tmp_path / "profiles").read_text().splitlines()
    assert len(man) == kept + 1 and "base profile filename = profile" in man[0]
    lines = (tmp_path / ("profile%d" % models[-1])).read_text().split("\n")
    assert lines[3].split()[:3] == ["model", "time", "core_mass"] and int(lines[4].split()[0]) == models[-1]
    assert lines[8].split()[0] == "zone" and lines[8].split()[-1] == "K" and len(lines[8]) == 20 * 15
    rows = [l for l in lines[9:] if l.strip()]
    assert len(rows) == nz and all(len(r) == 20 * 15 for r in rows)
    assert float(rows[0].split()[7]) == pytest.approx(q["temperature"][0], rel=1e-6)
    s.free()


def test_device_stack_limit_is_left_alone(nscool):
    """The library does not touch the process-wide cudaLimitStackSize (round 1 raised it to 8 KB, 2.4 GB of HBM, to
    steady its fused cell kernel; the round-2 kernels do not need it -- capi.cu ensure_stack_limit), and the tables do
    not depend on the window NCCL leaves behind (1872 B)."""
    import ctypes
    cu = ctypes.CDLL("libcuda.so.1")

    def limit():
        v = ctypes.c_size_t(0)
        assert cu.cuCtxGetLimit(ctypes.byref(v), 0) == 0   # CU_LIMIT_STACK_SIZE
        return v.value

    from dstar_b200 import workloads
    before = limit()
    s = workloads.basic_run(resolution=0.4)[0]
    s.create_model()
    assert limit() == before
    ref = {k: s.array(k).copy() for k in TABS}
    assert cu.cuCtxSetLimit(0, ctypes.c_size_t(1872)) == 0      # what NCCL does when it creates a communicator
    s2 = workloads.basic_run(resolution=0.4)[0]
    s2.create_model()
    assert limit() == 1872
    for k in TABS:
        assert np.array_equal(bits(s2.array(k)), bits(ref[k])), k
    assert cu.cuCtxSetLimit(0, ctypes.c_size_t(before)) == 0


def test_run_batch_driver_matches_create_and_evolve(nscool):
    """dstar_NScool_run_batch (the C-level sweep driver, one host thread per GPU): on one device it is create_models +
    evolve_models; on two devices -- when the box has them -- each device takes a contiguous block of the stars and
    every monitor, table and solver counter is bit-identical to the one-device run."""
    ds = nscool
    Tcs = [0.8e8, 1.0e8, 1.3e8, 1.7e8, 2.2e8]
    ref = [_star(ds, 0.3, core_temperature=t) for t in Tcs]
    ds.create_models(ref)
    ds.evolve_models(ref)
    import torch
    n_dev = torch.cuda.device_count()
    for n_gpu in ([1, 2] if n_dev >= 2 else [1]):
        run = [_star(ds, 0.3, core_temperature=t) for t in Tcs]
        tm, Tm, Qm, st = ds.run_batch(run, n_gpu=n_gpu)
        assert np.all(st == 0)
        assert Tm.shape == (2, len(Tcs))
        for i, (a, b) in enumerate(zip(ref, run)):
            for x, y in zip(a.monitors(), b.monitors()):
                assert np.array_equal(bits(x), bits(y)), (n_gpu, i)
            assert np.array_equal(bits(Tm[:, i]), bits(a.monitors()[1]))
            for name in TABS + ("lnT",):
                assert np.array_equal(bits(a.array(name)), bits(b.array(name))), (n_gpu, i, name)
            np.testing.assert_array_equal(a.counters(), b.counters())
        for s in run:
            s.free()
    for s in ref:
        s.free()


def test_two_batches_in_flight_on_one_device(nscool):
    """Two kernel-level contexts on the same device, one host thread each, stepping their own batch at the same time (the
    usage bench.py's `e2e.two_batches_in_flight` times and INTEGRATION.md section 2 recommends): every table and every
    monitor equals the one-batch-at-a-time result bit for bit, whichever way the two streams interleave."""
    import threading
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    ds = nscool
    W = bench.prepare_workload(ds, "custom_Tc_Qimp", 16, 0.2, 0)
    gaps = ("ns", "sfb03", "t72")
    ctxs = [ds.Context(0, gaps=gaps), ds.Context(0, gaps=gaps)]
    batches = [ds.Batch(c, W["zoff"], W["H"]["ncharged"]) for c in ctxs]
    try:
        def step(b):
            bench.upload_all(b, W)
            b.micro_tables(W["mc"], 3)
            b.evolve(W["ec"], 0)
            r = b.download_results(full=False)
            t = b.download_tables()
            return {k: np.array(r[k]).copy() for k in ("Teff_monitor", "t_monitor", "Qb_monitor", "counters", "idid")}, \
                   {k: np.array(t[k]).copy() for k in ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK")}
        ref_r, ref_t = step(batches[0])
        assert np.all(ref_r["idid"] >= 0)
        out, errs = [[], []], []

        def worker(i):
            try:
                for _ in range(4):
                    out[i].append(step(batches[i]))
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errs, errs
        for i in range(2):
            assert len(out[i]) == 4
            for r, t in out[i]:
                for k in ref_r:
                    np.testing.assert_array_equal(r[k], ref_r[k])
                for k in ref_t:
                    np.testing.assert_array_equal(t[k], ref_t[k])
    finally:
        for b in batches:
            b.close()
        for c in ctxs:
            c.close()
