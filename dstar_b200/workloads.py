"""Synthetic workloads of BASELINE.json `configs`, seed 20261019.

Each builder returns a list of dstar_b200.NScool slots with their controls set exactly as the
corresponding example inlist sets them (examples/<name>/inlist), varied per star as the sweep /
MCMC driver would -- including the outer boundary: no example inlist sets atm_model, so all of them run
the default 'bc09' envelope (NScool/defaults/controls.defaults:149), whose table the device generates.
"""
import numpy as np

from .api import NScool

SEED = 20261019

_COMMON = dict(atm_model="bc09", maximum_number_of_models=10000, maximum_timestep=0.0, integration_tolerance=1.0e-4,
               min_lg_temperature_integration=7.0, max_lg_temperature_integration=9.5, core_mass=1.6,
               core_radius=11.0, lgPcrust_bot=32.5, lgPcrust_top=27.0)


def basic_run(resolution=0.05):
    """configs[0]: examples/basic_run/inlist."""
    return [NScool(None, **_COMMON, target_resolution_lnP=resolution, fix_core_temperature=True,
                   core_temperature=1.4e8, fix_atmosphere_temperature_when_accreting=True,
                   atmosphere_temperature_when_accreting=2.4e8, number_epochs=2,
                   basic_epoch_Mdots=[1.0e17, 0.0], basic_epoch_boundaries=[0.0, 50.0, 1050.0],
                   turn_on_extra_heating=True, Q_heating_shallow=0.1, lgP_min_heating_shallow=27.5,
                   lgP_max_heating_shallow=28.5, turn_on_shell_Urca=False, which_neutron_1S0_gap="sfb03",
                   lg_atm_light_element_column=8.0, fix_Qimp=True, Qimp=1.0)]


def custom_Tc_Qimp(n_star=256, resolution=0.05, offset=0):
    """configs[1]: examples/custom_Tc_Qimp -- (core Tc, pasta Qimp) grid, hooks alt_Qimp / alt_sf, 9 epochs.
    `offset` shifts the grid so that different ranks get different stars (weak scaling)."""
    side = int(round(np.sqrt(n_star)))
    assert side * side == n_star, "n_star must be a square"
    Tc = np.logspace(np.log10(3.0e7), np.log10(3.0e8), side)
    Q = np.logspace(0.0, 2.0, side)
    stars = []
    for i in range(side):
        for j in range(side):
            # rank offset: a tiny deterministic shift of Tc keeps every star distinct across ranks
            s = NScool(None, **_COMMON, target_resolution_lnP=resolution, fix_core_temperature=True,
                       core_temperature=float(Tc[i] * (1.0 + 1.0e-3 * offset)),
                       fix_atmosphere_temperature_when_accreting=False,
                       atmosphere_temperature_when_accreting=2.4e8, number_epochs=9,
                       basic_epoch_Mdots=[1.0e17] + [0.0] * 8,
                       basic_epoch_boundaries=[-4383.0, 0.0, 65.1, 235.7, 751.6, 929.5, 1500.5, 1570.4, 1595.4, 3039.7],
                       use_other_set_Qimp=True, use_other_sf_critical_temperatures=True,
                       extra_real_controls=[4.0e13, float(Q[j])], turn_on_extra_heating=True,
                       Q_heating_shallow=1.7, lgP_min_heating_shallow=27.0, lgP_max_heating_shallow=28.0,
                       which_neutron_1S0_gap="sfb03", lg_atm_light_element_column=4.0, fix_Qimp=True, Qimp=0.0)
            s.use_example_hooks()
            stars.append(s)
    return stars


def fit_lightcurve(n_star=1024, resolution=0.05, offset=0):
    """configs[2]: examples/fit_lightcurve MCMC batch -- walkers over (Mdot, Q_heating_shallow, Qimp), 7 epochs."""
    rng = np.random.default_rng(SEED + 2 + 1000 * offset)
    Mdot = 10.0 ** rng.uniform(np.log10(3e16), np.log10(3e17), n_star)
    Qh = rng.uniform(0.0, 3.0, n_star)
    Qimp = 10.0 ** rng.uniform(np.log10(0.1), np.log10(30.0), n_star)
    stars = []
    for k in range(n_star):
        stars.append(NScool(None, **_COMMON, target_resolution_lnP=resolution, fix_core_temperature=True,
                            core_temperature=9.2e7, fix_atmosphere_temperature_when_accreting=False,
                            number_epochs=7, basic_epoch_Mdots=[float(Mdot[k])] + [0.0] * 6,
                            basic_epoch_boundaries=[-1000.0, 0.0, 50.0, 100.0, 500.0, 1000.0, 2000.0, 5000.0],
                            turn_on_extra_heating=True, Q_heating_shallow=float(Qh[k]),
                            lgP_min_heating_shallow=27.0, lgP_max_heating_shallow=28.0,
                            which_neutron_1S0_gap="sfb03", lg_atm_light_element_column=4.0, fix_Qimp=True,
                            Qimp=float(Qimp[k])))
    return stars


def accretion(n_star=4096, resolution=0.05, offset=0):
    """configs[3]: examples/accretion -- 2 cycles x 5 epochs (accretion_history), neutron + superfluid-phonon
    conductivity and shell Urca on."""
    rng = np.random.default_rng(SEED + 3 + 1000 * offset)
    scale = 10.0 ** rng.uniform(17.0, 18.0, n_star)
    Qimp = 10.0 ** rng.uniform(0.0, 2.0, n_star)
    t = np.array([-10000.0, 0.0, 10.0, 100.0, 1000.0, 10000.0])
    md = np.array([0.1, 0.0, 0.0, 0.0, 0.0])
    bounds = np.concatenate([t[:-1], t[:-1] + (t[-1] - t[0]), [t[-1] + (t[-1] - t[0])]])
    stars = []
    for k in range(n_star):
        stars.append(NScool(None, **_COMMON, target_resolution_lnP=resolution, fix_core_temperature=True,
                            core_temperature=3.0e8, fix_atmosphere_temperature_when_accreting=True,
                            atmosphere_temperature_when_accreting=5.0e8, number_epochs=10,
                            basic_epoch_Mdots=list(np.tile(md, 2) * scale[k]), basic_epoch_boundaries=list(bounds),
                            use_neutron_conductivity=True, use_superfluid_phonon_conductivity=True,
                            turn_on_shell_Urca=True, turn_on_extra_heating=True, Q_heating_shallow=1.0,
                            which_neutron_1S0_gap="gc", fix_Qimp=True, Qimp=float(Qimp[k])))
    return stars


def int16_demo(n_star=1024, resolution=0.05, offset=0):
    """configs[4]: examples/INT-16-2b-demo/inlist-Q4-H1.7 -- crust composition from an abuntime network-output table
    (synthetic file of the same format, tools/make_abuntime.py: 63 nuclides + neutrons per zone, blended into HZ90
    above lgP 30.6), 10 epochs; Qimp and the shallow heating vary per star."""
    rng = np.random.default_rng(SEED + 4 + 1000 * offset)
    Qimp = 10.0 ** rng.uniform(0.0, 2.0, n_star)
    Qh = rng.uniform(0.0, 3.0, n_star)
    common = dict(_COMMON)
    common["lgPcrust_bot"] = 32.9
    stars = []
    for k in range(n_star):
        stars.append(NScool(None, **common, target_resolution_lnP=resolution, fix_core_temperature=True,
                            core_temperature=9.2e7, fix_atmosphere_temperature_when_accreting=False, number_epochs=10,
                            basic_epoch_Mdots=[1.0e17] + [0.0] * 9,
                            basic_epoch_boundaries=[-4383.0, 0.0, 1.0, 3.0, 10.0, 30.0, 100.0, 300.0, 1000.0, 3000.0, 6000.0],
                            turn_on_extra_heating=True, Q_heating_shallow=float(Qh[k]), lgP_min_heating_shallow=27.0,
                            lgP_max_heating_shallow=28.0, turn_on_shell_Urca=False, which_neutron_1S0_gap="sfb03",
                            lg_atm_light_element_column=4.0, fix_Qimp=True, Qimp=float(Qimp[k]),
                            crust_composition="int16_synth"))
    return stars


BUILDERS = {"int16_demo": int16_demo, "basic_run": basic_run, "custom_Tc_Qimp": custom_Tc_Qimp, "fit_lightcurve": fit_lightcurve,
            "accretion": accretion}
