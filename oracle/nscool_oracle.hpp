// ORACLE (test infrastructure): NScool coefficient pass and implicit thermal step.
// POD structs are laid out for ctypes (tests/oracle_lib.py mirrors them field by field).
#pragma once
#include <cstdint>

#include "dstar_oracle.hpp"

namespace dso {

constexpr int number_table_pts = 128;  // NScool_create_model.f:16
constexpr double lgT_tab_min = 7.0, lgT_tab_max = 10.0;

// Inputs of do_setup_crust_transport (NScool_create_model.f:289-444) for ONE star.
struct MicroTablesIn {
    int32_t nz, ncharged;
    const double *rho, *rho_bar, *P, *P_bar, *dm;  // (nz)
    const double *ionic, *ionic_bar;               // (11, nz)
    const double *Yion, *Yion_bar;                 // (ncharged, nz)
    const double *Zion, *Aion;                     // (ncharged)
    const double *Tc_cell, *Tc_face;  // (3, nz) hook-evaluated critical temperatures, or null -> gap tables
    double eos_ctrl[3];               // Gamma_melt, rsi_melt, Ythresh
    double cond_ctrl[6];              // include_neutrons, include_sf_phonons, ee_fmla, eQ_fmla, rad_off, rad_on
    int32_t nu_channels[5];           // pair, photo, plasma, brems, pbf
    int32_t turn_on_shell_Urca;
    double shell_Urca_luminosity_coeff, lgP_shell_Urca;
    int32_t do_cells, do_faces;       // which of the two loops to run
    double rho_bar1_override;         // used for face 1 when do_cells == 0 (<=0: use rho_bar[0])
};
int micro_tables(const HelmTable& h, const MicroTablesIn& in, double* tab_lnT, double* tab_lnCp,
                 double* tab_lnGamma, double* tab_lnEnu, double* tab_lnK, double* rho_bar1);
void set_micro_sf_tables(const SfTables* sf);

// dStar_crust_get_composition_info (dStar_crust_lib.f:109-147) on a prebuilt table
int crust_composition_info(int nv, int nisos, const double* lgP_tab, const double* Ycoef,
                           const double* Ziso, const double* Aiso, int nz, const double* lgP,
                           double* ionic, double* Yion, int* ncharged);

// Inputs of do_integrate_crust (NScool_evolve.f:11-136) for ONE star, ONE epoch.
struct EvolveIn {
    int32_t nz, n_tab;
    const double* tab_lnT;                                      // (n_tab)
    const double *tab_lnK, *tab_lnCp, *tab_lnGamma, *tab_lnEnu;  // (4*n_tab, nz)
    const double *dm, *dm_bar, *area, *rho_bar, *ePhi, *e2Phi, *ePhi_bar, *e2Phi_bar, *P;  // (nz)
    int32_t atm_nv;
    const double *atm_lgTb, *atm_lgTeff, *atm_lgflux;  // (nv), (4*nv), (4*nv)
    // heating windows {lgPmin, lgPmax, Q} outer, inner, shallow (NScool_evolve.f:526-568)
    double heat[9];
    int32_t turn_on_extra_heating;
    const double* enuc_override;  // (nz) host-evaluated other_set_heating result or null
    int32_t fix_core_temperature, make_inner_boundary_insulating;
    int32_t fix_atmosphere_temperature_when_accreting;
    double atmosphere_temperature_when_accreting;
    double Mdot, epoch_start_time, epoch_duration;
    double integration_tolerance, min_lg_T, max_lg_T, maximum_timestep;
    int32_t maximum_number_of_models;
    int32_t starting_number_for_profile, suppress_first_step_output;
    int32_t reference_cost_quirks;  // 1: redo the whole-array exps per zone as NScool_evolve.f:504-507 does
};
struct EvolveState {
    double* lnT;  // (nz) in/out
    double dt, tsec;
    int32_t model;
    double Teff, Lsurf, dlnLsdlnT;
    int32_t idid;
    int32_t counters[7];  // nfcn, njac, nstep, naccpt, nrejct, ndec, nsol
};
struct EvolveTrace {  // one row per solout call (history columns, NScool_history.f:84-86)
    int32_t capacity, count;
    int32_t* model;
    double *tsec, *dt, *Teff, *Lsurf, *Lnu, *Lnuc;
    double* lnT_rows;  // optional (capacity, nz) snapshots, may be null
};
int evolve_epoch(const EvolveIn& in, EvolveState& st, EvolveTrace* trace);
int evolve_rhs_jac(const EvolveIn& in, const double* lnT, double* f, double* lo, double* di, double* up, double* surf);

// RODASP order/self-consistency test problems (tests/test_oracle_integrator.py)
int rodasp_selftest(int problem, double tend, double rtol, double atol, double h0, double* y,
                    int* counters);

}  // namespace dso
