"""Run the CPU oracle on the inputs of a product-built star (tests / smoke only)."""
import numpy as np

from fixtures import load_gap_file

JULIAN_DAY = 86400.0
AMU = 1.0 / 6.02214076e23
MEV = 1.602176634e-6


def load_oracle_gaps(o, p1="ns", n1="gc", n3="t72", scale=(1.0, 1.0, 1.0)):
    for typ, (name, code) in enumerate(((p1, "p1"), (n1, "n1"), (n3, "n3"))):
        kF, Tc, kmin, kmax = load_gap_file(name, code)
        o.sf_load(typ, kF, Tc, kmin, kmax)
    o.sf_scale(scale)


def oracle_micro_tables(o, s, gaps=("ns", "sfb03", "t72"), cond_ctrl=(0, 0, 1, 1, 10.0, 9.0), Tc_cell=None,
                        Tc_face=None, shell_Urca=(0, 1.0, 28.5)):
    """s: dstar_b200.NScool after create_model.  Feeds the oracle the star's pre-fix-up inputs."""
    load_oracle_gaps(o, *gaps)
    nz = s.get_int("nz"); nch = s.get_int("ncharged")
    rho_bar = s.array("rho_bar").copy()
    rho_bar[0] = 0.0   # undefined before the coefficient pass (create_model.f:231-232)
    return o.micro_tables(nz, nch, s.array("rho"), rho_bar, s.array("P"), s.array("P_bar"), s.array("dm"),
                          s.array("ionic"), s.array("ionic_bar"), s.array("Yion"), s.array("Yion_bar"),
                          s.array("Zion"), s.array("Aion"), Tc_cell=Tc_cell, Tc_face=Tc_face, cond_ctrl=cond_ctrl,
                          shell_Urca=shell_Urca)


def evolve_in_kwargs(s, epoch, tabs=None, ctrl=None, heat=None, T_atm=None):
    """EvolveIn fields of one epoch of star s (same defaults as oracle_evolve_star)."""
    nz = s.get_int("nz")
    t = tabs or {k: s.array(k) for k in ("tab_lnT", "tab_lnK", "tab_lnCp", "tab_lnGamma", "tab_lnEnu")}
    if "tab_lnT" not in t:
        t["tab_lnT"] = s.array("tab_lnT")
    eb = s.array("epoch_boundaries"); em = s.array("epoch_Mdots")
    kw = dict(nz=nz, n_tab=128, tab_lnT=t["tab_lnT"], tab_lnK=t["tab_lnK"], tab_lnCp=t["tab_lnCp"],
              tab_lnGamma=t["tab_lnGamma"], tab_lnEnu=t["tab_lnEnu"], dm=s.array("dm"), dm_bar=s.array("dm_bar"),
              area=s.array("area"), rho_bar=s.array("rho_bar"), ePhi=s.array("ePhi"), e2Phi=s.array("e2Phi"),
              ePhi_bar=s.array("ePhi_bar"), e2Phi_bar=s.array("e2Phi_bar"), P=s.array("P"), atm_nv=256,
              atm_lgTb=s.array("atm_lgTb"), atm_lgTeff=s.array("atm_lgTeff"), atm_lgflux=s.array("atm_lgflux"),
              heat=heat or [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 3.0], turn_on_extra_heating=0,
              fix_core_temperature=1, make_inner_boundary_insulating=0,
              fix_atmosphere_temperature_when_accreting=1,
              atmosphere_temperature_when_accreting=T_atm or 1.0e8, integration_tolerance=1.0e-4, min_lg_T=7.0,
              max_lg_T=9.5, maximum_timestep=0.0, maximum_number_of_models=10000, starting_number_for_profile=0,
              suppress_first_step_output=0, reference_cost_quirks=0)
    if ctrl:
        kw.update(ctrl)
    kw["Mdot"] = float(em[epoch]); kw["epoch_start_time"] = float(eb[epoch])
    kw["epoch_duration"] = float(eb[epoch + 1] - eb[epoch])
    return kw


def oracle_evolve_star(o, s, lnT0, tabs=None, ctrl=None, heat=None, T_atm=None, trace_cap=0, quirks=0):
    """NScool_evolve_model's epoch loop (NScool_lib.f:68-90) on the oracle, from the star's own tables
    (or `tabs`).  ctrl: dict overriding EvolveIn scalar fields."""
    nz = s.get_int("nz")
    t = tabs or {k: s.array(k) for k in ("tab_lnT", "tab_lnK", "tab_lnCp", "tab_lnGamma", "tab_lnEnu")}
    if "tab_lnT" not in t:
        t["tab_lnT"] = s.array("tab_lnT")
    eb = s.array("epoch_boundaries"); em = s.array("epoch_Mdots")
    base = dict(nz=nz, n_tab=128, tab_lnT=t["tab_lnT"], tab_lnK=t["tab_lnK"], tab_lnCp=t["tab_lnCp"],
                tab_lnGamma=t["tab_lnGamma"], tab_lnEnu=t["tab_lnEnu"], dm=s.array("dm"), dm_bar=s.array("dm_bar"),
                area=s.array("area"), rho_bar=s.array("rho_bar"), ePhi=s.array("ePhi"), e2Phi=s.array("e2Phi"),
                ePhi_bar=s.array("ePhi_bar"), e2Phi_bar=s.array("e2Phi_bar"), P=s.array("P"), atm_nv=256,
                atm_lgTb=s.array("atm_lgTb"), atm_lgTeff=s.array("atm_lgTeff"), atm_lgflux=s.array("atm_lgflux"),
                heat=heat or [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.0, 28.0, 3.0], turn_on_extra_heating=0,
                fix_core_temperature=1, make_inner_boundary_insulating=0,
                fix_atmosphere_temperature_when_accreting=1,
                atmosphere_temperature_when_accreting=T_atm or 1.0e8, integration_tolerance=1.0e-4, min_lg_T=7.0,
                max_lg_T=9.5, maximum_timestep=0.0, maximum_number_of_models=10000, starting_number_for_profile=0,
                suppress_first_step_output=0, reference_cost_quirks=quirks)
    if ctrl:
        base.update(ctrl)
    ePhi1 = s.array("ePhi")[0]
    lnT = np.array(lnT0, dtype=float).copy()
    dt = 0.0
    model = 0
    out = {"t_monitor": [], "Teff_monitor": [], "Qb_monitor": [], "counters": np.zeros(7, dtype=int), "trace": []}
    start0 = base["starting_number_for_profile"]
    for ie in range(len(em)):
        kw = dict(base)
        kw["Mdot"] = float(em[ie]); kw["epoch_start_time"] = float(eb[ie]); kw["epoch_duration"] = float(eb[ie + 1] - eb[ie])
        if ie > 0:
            kw["starting_number_for_profile"] = model if base["suppress_first_step_output"] else model + 1
        else:
            kw["starting_number_for_profile"] = start0
        r = o.evolve_epoch(kw, lnT, dt, trace_cap=trace_cap)
        assert r["idid"] >= 0, r["idid"]
        lnT = r["lnT"]; dt = r["dt"]; model = r["model"]
        out["t_monitor"].append(r["tsec"] / JULIAN_DAY)
        out["Teff_monitor"].append(r["Teff"] * ePhi1)
        out["Qb_monitor"].append(r["Lsurf"] / em[ie] * AMU / MEV if em[ie] > 0 else 0.0)
        out["counters"] += r["counters"]
        if trace_cap:
            out["trace"].append(r["trace"])
    for k in ("t_monitor", "Teff_monitor", "Qb_monitor"):
        out[k] = np.array(out[k])
    out["lnT"] = lnT
    return out
