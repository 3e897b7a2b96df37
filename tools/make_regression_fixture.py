"""Generates tests/golden/basic_run_small.npz on a GPU box: the inputs of one small basic_run star (structure,
composition, atmosphere table, epochs, initial ln T) as the product's host layer built them, and the product's
outputs for them (coefficient tables, monitors, solver counters, final ln T).  tests/test_regression_fixture.py
replays the inputs through the CPU oracle without a GPU (oracle == fixture to 1e-10 / 1e-6) and through the CUDA
path on a GPU (bitwise for the tables).  Usage (GPU box): python tools/make_regression_fixture.py OUT.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ("rho", "rho_bar", "P", "P_bar", "dm", "ionic", "ionic_bar", "Yion", "Yion_bar", "Zion", "Aion", "tab_lnT",
         "dm_bar", "area", "ePhi", "e2Phi", "ePhi_bar", "e2Phi_bar", "atm_lgTb", "atm_lgTeff", "atm_lgflux",
         "epoch_boundaries", "epoch_Mdots")
CONTROLS = dict(atm_model="pcy97", target_resolution_lnP=0.6, number_epochs=2, basic_epoch_Mdots=[1.0e17, 0.0],
                basic_epoch_boundaries=[0.0, 50.0, 1050.0], core_temperature=1.4e8,
                atmosphere_temperature_when_accreting=2.4e8, turn_on_extra_heating=True, Q_heating_shallow=0.1,
                lgP_min_heating_shallow=27.5, lgP_max_heating_shallow=28.5, which_neutron_1S0_gap="sfb03",
                fix_Qimp=True, Qimp=1.0)


def main(out):
    import dstar_b200 as ds
    ds.NScool_init()
    s = ds.NScool(None, **CONTROLS)
    s.create_model()
    d = {k: s.array(k).copy() for k in NAMES}
    d["rho_bar_in"] = d["rho_bar"].copy()          # after the rho_bar(1) fix-up (the oracle helper zeroes element 0)
    d["nz"] = np.int64(s.get_int("nz")); d["ncharged"] = np.int64(s.get_int("ncharged"))
    d["lnT0"] = s.array("lnT").copy()
    for k in ("tab_lnCp", "tab_lnGamma", "tab_lnEnu", "tab_lnK"):
        d["gpu_" + k] = s.array(k).copy()
    s.evolve_model()
    t, Teff, Qb = s.monitors()
    d["gpu_t_monitor"] = t; d["gpu_Teff_monitor"] = Teff; d["gpu_Qb_monitor"] = Qb
    d["gpu_counters"] = np.asarray(s.counters(), dtype=np.int64); d["gpu_lnT"] = s.array("lnT").copy()
    np.savez_compressed(out, **d)
    print("wrote", out, "nz =", int(d["nz"]), "Teff_monitor =", Teff)
    ds.NScool_shutdown()


if __name__ == "__main__":
    main(sys.argv[1])
