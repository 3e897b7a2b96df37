"""Generates tests/golden/output_sample/ on a GPU box: LOGS-style output of one small basic_run star written by the
product's writers (history.data, profiles + manifest, terminal log) together with the raw arrays they were written
from (arrays.npz).  tests/test_output_files.py parses the committed files -- with the reference's own
tools/reader.py when /root/reference is present, otherwise with a parser of the same layout rules -- and compares
them with the arrays.  Usage (GPU box): python tools/make_output_fixture.py tests/golden/output_sample
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONTROLS = dict(atm_model="bc09", target_resolution_lnP=0.5, number_epochs=2, basic_epoch_Mdots=[1.0e17, 0.0],
                basic_epoch_boundaries=[0.0, 50.0, 1050.0], core_temperature=1.4e8,
                atmosphere_temperature_when_accreting=2.4e8, turn_on_extra_heating=True, Q_heating_shallow=0.1,
                lgP_min_heating_shallow=27.5, lgP_max_heating_shallow=28.5, which_neutron_1S0_gap="sfb03",
                fix_Qimp=True, Qimp=1.0, write_interval_for_history=1, write_interval_for_profile=10,
                write_interval_for_terminal=2, write_interval_for_terminal_header=5)


def main(out):
    import dstar_b200 as ds
    os.makedirs(out, exist_ok=True)
    ds.NScool_init()
    s = ds.NScool(None, **CONTROLS)
    s.create_model()
    s.evolve_model()
    assert s.write_history(os.path.join(out, "history.data")) == 0
    assert s.write_terminal(os.path.join(out, "terminal.txt")) == 0
    s.write_profiles(out)
    h = s.history(); term = s.terminal()
    d = {"h_" + k: v for k, v in h.items()}
    d.update({"t_" + k: v for k, v in term.items()})
    kept, made = s.profile_count()
    d["n_profiles"] = np.int64(kept)
    for i in range(kept):
        pr = s.profile(i)
        d["p%d_model" % i] = np.int64(pr["model"])
        d["p%d_time" % i] = np.float64(pr["time"])
        for c in ("temperature", "luminosity", "pressure", "density", "cp", "eps_nu", "K", "Gamma"):
            d["p%d_%s" % (i, c)] = pr[c]
    for k in ("grav", "Mtotal", "Rtotal", "Mcore", "Rcore", "Tcore"):
        d[k] = np.float64(s.get_real(k))
    np.savez_compressed(os.path.join(out, "arrays.npz"), **d)
    print("wrote", out, "history rows", len(h["model"]), "profiles", kept, sorted(os.listdir(out)))
    ds.NScool_shutdown()


if __name__ == "__main__":
    main(sys.argv[1])
