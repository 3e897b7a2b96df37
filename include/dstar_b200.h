/* dstar_b200 -- C ABI of the B200-native dStar hot path.
 *
 * The reference (nworbde/dStar) has no FFI layer: its "operator API" is the Fortran module API of
 * NScool_lib (NScool/public/NScool_lib.f:6-91) over type NScool_info (NScool/public/NScool_def.f:33-157).
 * This header is what a thin ISO_C_BINDING module (fortran/dstar_b200_bindc.f90)
 * binds so that NScool_create_model / NScool_evolve_model delegate their two hot loops to the GPU:
 *
 *   seam A  body of do_setup_crust_transport  NScool/private/NScool_create_model.f:340-441
 *           -> dstar_batch_micro_tables / dstar_micro_tables
 *   seam B  the isolve(...) call              NScool/private/NScool_evolve.f:111-117, driven by the
 *           epoch loop of NScool_evolve_model NScool/public/NScool_lib.f:68-90
 *           -> dstar_batch_evolve / dstar_evolve
 *
 * Conventions kept from the reference: every call returns an integer status, 0 = ok, < 0 = failure
 * (the Fortran `ierr`); arrays are column-major REAL(8) / INTEGER(4) with the Fortran shapes quoted
 * in the comments; the caller owns all host buffers; the library owns device memory behind opaque
 * handles.  One host thread per GPU.  Nothing here falls back to the CPU: every entry point fails
 * with DSTAR_ERR_CUDA if no sm_100-class device is usable.
 */
#ifndef DSTAR_B200_H
#define DSTAR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSTAR_OK 0
#define DSTAR_ERR_CUDA -101     /* no device / CUDA failure (message via dstar_last_error) */
#define DSTAR_ERR_ARG -102      /* bad argument */
#define DSTAR_ERR_STATE -103    /* call order violated (table not loaded, batch not uploaded ...) */
#define DSTAR_NUM_EOS_RESULTS 15 /* dStar_eos/public/dStar_eos_def.f:36 */
#define DSTAR_NTAB 128           /* NScool/private/NScool_create_model.f:16 */

typedef struct dstar_ctx dstar_ctx;     /* replaces NScool_init / *_startup / *_shutdown state */
typedef struct dstar_batch dstar_batch; /* a set of stars (NScool_info slots) resident in HBM */

/* ---- context: device + read-only tables (replaces dStar_eos_startup, sf_load_gaps, ...) ---- */
int dstar_ctx_create(int device, dstar_ctx** ctx);
void dstar_ctx_destroy(dstar_ctx* ctx);
const char* dstar_last_error(void);
/* HELM table: 21 arrays (imax,jmax) in the order helm_alloc.f:176-196 reads them
 * (f fd ft fdd ftt fdt fddt fdtt fddtt dpdf dpdfd dpdft dpdfdt ef efd eft efdt xf xfd xft xfdt);
 * replaces read_helm_table (dStar_eos/private/helm_alloc.f:120-330). */
int dstar_ctx_set_helm(dstar_ctx* ctx, int imax, int jmax, const double* const arrays[21]);
/* gap table `type` (0 = p 1S0, 1 = n 1S0, 2 = n 3P2): kF(nv) and the (4,nv) monotone-cubic
 * coefficients load_sf_table builds (superfluid/private/load_sf.f:60-69). */
int dstar_ctx_set_gaps(dstar_ctx* ctx, int type, int nv, double kF_min, double kF_max, const double* kF,
                       const double* f4);
int dstar_ctx_set_sf_scale(dstar_ctx* ctx, const double scale[3]); /* sf_scale, superfluid_def.f:29 */

/* ---- batched point evaluators (one call = n independent evaluations, host buffers) ---- */
/* eval_crust_eos (dStar_eos/public/dStar_eos_lib.f:123): ionic(11,n), Yion(ncharged,n), Tcs(3,n),
 * chi_in(n) or NULL (use_default_nuclear_size), eos_ctrl = {Gamma_melt, rsi_melt, Ythresh} or NULL;
 * res(15,n), phase(n), chi_out(n), ierr(n). */
int dstar_eval_crust_eos(dstar_ctx* ctx, int n, const double* rho, const double* T, const double* ionic,
                         int ncharged, const double* Zion, const double* Aion, const double* Yion,
                         const double* Tcs, const double* chi_in, const double* eos_ctrl, double* res,
                         int32_t* phase, double* chi_out, int32_t* ierr);
/* same, with the optional `components` argument (dStar_eos_lib.f:260-296): components(7,4,n) =
 * {P,E,S,F,Cv,dPdlnRho,dPdlnT} x {electron+exchange, ion, neutron, radiation} (dStar_eos_def.f:38-56) */
int dstar_eval_crust_eos_components(dstar_ctx* ctx, int n, const double* rho, const double* T,
                                    const double* ionic, int ncharged, const double* Zion, const double* Aion,
                                    const double* Yion, const double* Tcs, const double* chi_in,
                                    const double* eos_ctrl, double* res, int32_t* phase, double* chi_out,
                                    int32_t* ierr, double* components);
/* get_crust_neutrino_emissivity (neutrino/public/neutrino_lib.f:16): channels[5] = pair, photo, plasma,
 * brems, pbf; eps(6,n) = total, pair, photo, plasma, brems, pbf. */
int dstar_crust_neutrino(dstar_ctx* ctx, int n, const double* rho, const double* T, const double* ionic,
                         const double* chi, const double* Tcn, const int32_t channels[5], double* eps);
/* get_thermal_conductivity (conductivity/public/conductivity_lib.f:108): cond_ctrl = {include_neutrons,
 * include_superfluid_phonons, ee_fmla(1 SY06|2 PCY), eQ_fmla(1 Potekhin|2 Page), rad_full_off_lgrho,
 * rad_full_on_lgrho}; K(10,n) in the field order of conductivity_def.f:16-27. */
int dstar_conductivity(dstar_ctx* ctx, int n, const double* rho, const double* T, const double* chi,
                       const double* Gamma, const double* eta, const double* mu_e, const double* ionic,
                       const double* Tns, const double cond_ctrl[6], double* K);
/* sf_get_results (superfluid/public/superfluid_lib.f:65): Tc(3,n). */
int dstar_sf_get_results(dstar_ctx* ctx, int n, const double* kFp, const double* kFn, double* Tc);
/* dStar_crust_get_composition_info (dStar_crust/public/dStar_crust_lib.f:109) on a device-resident
 * composition table: lgP_tab(nv), Ycoef(4*nv,nisos), Ziso/Aiso(nisos); lgP(nz) -> ionic(11,nz),
 * Yion(ncharged,nz); *ncharged is set. */
int dstar_crust_composition_info(dstar_ctx* ctx, int nv, int nisos, const double* lgP_tab,
                                 const double* Ycoef, const double* Ziso, const double* Aiso, int nz,
                                 const double* lgP, double* ionic, double* Yion, int32_t* ncharged);

/* do_get_bc09_Teff (dStar_atm/private/bc09.f:45-95; the default atm_model, NScool/defaults/controls.defaults:149)
 * for n_tab independent (grav, Plight, Pb) triples: lgTeff(nv) ascending knots in, lgTb(nv,n_tab) and
 * lgflux(nv,n_tab) out, ierr(n_tab) = the reference's ierr per table.  info: NULL or (4,nv,n_tab) = rho_ph, P_ph,
 * kappa_ph, number of EOS evaluations.  What do_generate_atm_table (dStar_atm_mod.f:46-106) then does with the
 * result (interp_pm in lgTb) is host work: dstar_NScool_create_model(s) in dstar_nscool.h. */
int dstar_atm_bc09_Teff(dstar_ctx* ctx, int n_tab, const double* grav, const double* Plight, const double* Pb,
                        int nv, const double* lgTeff, double* lgTb, double* lgflux, int32_t* ierr, double* info);

/* ---- controls shared by all stars of a batch (subset of NScool/public/NScool_controls.inc) ---- */
typedef struct dstar_micro_ctrl {
    double eos_gamma_melt_pt, eos_rsi_melt_pt, eos_nuclide_abundance_threshold; /* <= 0: defaults */
    int32_t use_neutron_conductivity, use_superfluid_phonon_conductivity;
    int32_t use_pcy_for_ee_scattering, use_page_for_eQ_scattering;
    double rad_full_off_lgrho, rad_full_on_lgrho; /* <= 0: defaults 10, 9 */
    int32_t use_crust_nu_pair, use_crust_nu_photo, use_crust_nu_plasma, use_crust_nu_bremsstrahlung,
        use_crust_nu_pbf;
    int32_t turn_on_shell_Urca;
    double shell_Urca_luminosity_coeff, lgP_shell_Urca;
} dstar_micro_ctrl;

typedef struct dstar_evolve_ctrl {
    int32_t fix_core_temperature, make_inner_boundary_insulating;
    int32_t fix_atmosphere_temperature_when_accreting;
    int32_t turn_on_extra_heating;
    int32_t suppress_first_step_output, starting_number_for_profile;
    int32_t maximum_number_of_models;
    double integration_tolerance, min_lg_temperature_integration, max_lg_temperature_integration;
    double maximum_timestep; /* seconds; 0 = epoch length */
} dstar_evolve_ctrl;

/* ---- batch: stars flattened zone-major; zone_offset(n_star+1) gives each star's zone range ---- */
int dstar_batch_create(dstar_ctx* ctx, int n_star, const int32_t* zone_offset, int ncharged, dstar_batch** b);
void dstar_batch_destroy(dstar_batch* b);
/* H2D of the structure arrays (type NScool_info, NScool_def.f:85-126).  All (total_zones) unless noted:
 * ionic/ionic_bar (11,total), Yion/Yion_bar (ncharged,total), Zion/Aion (ncharged).
 * Tc_cell/Tc_face (3,total): critical temperatures already evaluated by other_sf_get_results on the
 * host, or NULL to use the context's gap tables (sf_get_results).
 * The copies are issued on two streams (what only the face pass needs overlaps the cell pass) and this call and
 * dstar_batch_upload_evolve do not wait for them: PAGE-LOCKED host arrays must stay unchanged until the next
 * synchronising call (dstar_batch_micro_tables, dstar_batch_evolve, any download); pageable arrays (what a Fortran
 * caller passes) are staged by the driver before the call returns. */
int dstar_batch_upload_structure(dstar_batch* b, const double* rho, const double* rho_bar, const double* P,
                                 const double* P_bar, const double* dm, const double* ionic,
                                 const double* ionic_bar, const double* Yion, const double* Yion_bar,
                                 const double* Zion, const double* Aion, const double* Tc_cell,
                                 const double* Tc_face);
/* seam A on resident data.  which: 1 = cell loop (lnCp, lnGamma, lnEnu, rho_bar(1) fix-up,
 * create_model.f:344-403), 2 = face loop (lnK, :406-441), 3 = both. */
int dstar_batch_micro_tables(dstar_batch* b, const dstar_micro_ctrl* ctrl, int which);
/* Table sharing for sweeps (hard part 8): stars whose coefficient-pass inputs are identical -- same
 * crust, composition and gaps for the cell tables; additionally the same impurity parameter and face critical
 * temperatures for the face tables -- need one table set between them.  owner(n_star): owner[i] = index of the first
 * star of i's group (owner[i] == i for a star that owns its tables; owner[i] <= i); the following coefficient pass
 * evaluates only the owners' zones, the device keeps one table set per owner, and the evolve kernel reads a star's
 * coefficients from its owner's set.  which = 1: cell tables (lnCp, lnGamma, lnEnu and the rho_bar(1) fix-up), 2: face
 * tables (lnK).  owner == NULL ends the sharing.  Call before the dstar_batch_micro_tables pass it concerns; results
 * are bit-identical to the unshared pass because the inputs are.  (The reference has no counterpart: each NScool_info
 * slot tabulates for itself, NScool_create_model.f:340-441.) */
int dstar_batch_set_table_sharing(dstar_batch* b, int which, const int32_t* owner);
/* D2H: tab_lnT(128); tab_* as (4*128, total_zones) (bit layout of NScool_def.f:130-134 per star);
 * rho_bar1(n_star); status(n_star).  Any pointer may be NULL. */
int dstar_batch_download_tables(dstar_batch* b, double* tab_lnT, double* tab_lnCp, double* tab_lnGamma,
                                double* tab_lnEnu, double* tab_lnK, double* rho_bar1, int32_t* status);
/* replace face Tc / rho_bar(1)-dependent inputs between the two loops (hook path) */
int dstar_batch_get_rho_bar1(dstar_batch* b, double* rho_bar1);
int dstar_batch_set_Tc_face(dstar_batch* b, const double* Tc_face);
int dstar_batch_set_ionic_bar(dstar_batch* b, const double* ionic_bar); /* after other_set_Qimp */

/* H2D of what seam B needs beyond the tables: geometry/metric (total_zones each); atmosphere tables
 * lgTb(atm_nv,n_atm), lgTeff/lgflux (4*atm_nv,n_atm) as built by do_generate_atm_table
 * (dStar_atm/private/dStar_atm_mod.f:46-106), atm_index(n_star); heat(9,n_star) = lgP_min, lgP_max, Q of
 * the outer, inner and shallow heating windows; T_atm(n_star); lnT0(total_zones) initial ln T;
 * epoch_boundaries(n_epoch+1,n_star) [s], epoch_Mdots(n_epoch,n_star) [g/s];
 * enuc_per_Mdot(total_zones) or NULL: host-evaluated other_set_heating result divided by Mdot. */
int dstar_batch_upload_evolve(dstar_batch* b, const double* dm_bar, const double* area, const double* ePhi,
                              const double* e2Phi, const double* ePhi_bar, const double* e2Phi_bar,
                              int n_atm, int atm_nv, const double* atm_lgTb, const double* atm_lgTeff,
                              const double* atm_lgflux, const int32_t* atm_index, const double* heat,
                              const double* T_atm, const double* lnT0, int n_epoch,
                              const double* epoch_boundaries, const double* epoch_Mdots,
                              const double* enuc_per_Mdot);
/* optionally take the tables from the host instead of dstar_batch_micro_tables */
int dstar_batch_upload_tables(dstar_batch* b, const double* tab_lnT, const double* tab_lnCp,
                              const double* tab_lnGamma, const double* tab_lnEnu, const double* tab_lnK);
/* restore the evolving state (lnT, dt, tsec, model) to what dstar_batch_upload_evolve set, device to device */
int dstar_batch_reset_state(dstar_batch* b);
/* seam B: every star through all its epochs in ONE launch.  trace_cap > 0 keeps that many
 * history rows per star. */
int dstar_batch_evolve(dstar_batch* b, const dstar_evolve_ctrl* ctrl, int trace_cap);
/* test support, function level: get_derivatives (NScool/private/NScool_evolve.f:237-283) and the tridiagonal rows of
 * get_jacobian (:285-359) of epoch `epoch` at the batch's current ln T (as uploaded or as the last evolve left it),
 * computed by the device functions the evolve kernel is built from.  f (total_zones); jac (3, total_zones) =
 * d f_k / d lnT_{k-1}, d f_k / d lnT_k, d f_k / d lnT_{k+1}; surf (3, n_star) = Lsurf, dlnLs/dlnT, Teff.  May be NULL. */
int dstar_batch_eval_rhs(dstar_batch* b, const dstar_evolve_ctrl* ctrl, int epoch, double* f, double* jac, double* surf);
/* thread mapping of seam B: 0 = automatic (default), 1 = one CTA per star, one thread per zone, cyclic reduction through
 * shared memory (any nz <= 512), 2 = one warp per star, nz/32 zones per lane, partitioned solve with cyclic reduction
 * over warp shuffles (nz <= 320; larger stars fall back to 1).  Same formulas and driver, so the same light curves;
 * the CTA form is the faster one on B200 at every batch size measured (profiles/r02_notes.md) and is what 0 selects. */
int dstar_batch_set_evolve_mapping(dstar_batch* b, int mapping);
/* tuning knob of mapping 1: how many one-star CTAs the evolve kernel is compiled to keep on an SM (1..4; register
 * budget 255 / 128 / 80 / 64 per thread).  0 = automatic: 1 up to one star per SM, 2 up to two, ... -- fewer resident
 * stars run each star faster, more resident stars finish a large batch sooner.  Results do not depend on it. */
int dstar_batch_set_evolve_occupancy(dstar_batch* b, int ctas_per_sm);
/* D2H: monitors (n_epoch,n_star) as NScool_def.f:136-144; idid (n_epoch,n_star); counters (7,n_star) =
 * nfcn, njac, nstep, naccpt, nrejct, ndec, nsol (isolve iwork(14:20)); lnT(total_zones); Teff, Lsurf,
 * dt, tsec (n_star); model(n_star).  Any pointer may be NULL. */
int dstar_batch_download_results(dstar_batch* b, double* t_monitor, double* Teff_monitor, double* Qb_monitor,
                                 int32_t* idid, int32_t* counters, double* lnT, double* Teff, double* Lsurf,
                                 double* dt, double* tsec, int32_t* model);
/* D2H of the trace: count(n_star); the rest (trace_cap,n_star): model, tsec, dt, Teff, Lsurf, Lnu, Lnuc, Mdot */
int dstar_batch_download_trace(dstar_batch* b, int32_t* count, int32_t* model, double* tsec, double* dt,
                               double* Teff, double* Lsurf, double* Lnu, double* Lnuc, double* Mdot);

/* D2H of the extrema the terminal line prints (NScool/private/NScool_terminal.f:53-56), per trace row:
 * ext (4, trace_cap, n_star) = maxval(lnT), P(maxloc(lnT)), minval(lnT), P(minloc(T)) */
int dstar_batch_download_trace_extrema(dstar_batch* b, double* ext);

/* Profile capture (evaluate_timestep, NScool/private/NScool_evolve.f:227-233 -> do_write_profile,
 * NScool/private/NScool_profile.f:27-106): before dstar_batch_evolve, ask the kernel to keep ln T of every zone at
 * the models with mod(model, interval) == 0, at most `capacity` per star (interval or capacity 0: off).
 * download: count(n_star) = profiles the run produced (may exceed capacity; only the first `capacity` are kept),
 * model/tsec/Mdot (capacity, n_star), lnT (capacity, total_zones).  Every other profile column (L, Gamma, Cp,
 * eps_nu, K) follows from ln T through the coefficient tables on the host. */
int dstar_batch_set_profile_capture(dstar_batch* b, int interval, int capacity);
int dstar_batch_download_profiles(dstar_batch* b, int32_t* count, int32_t* model, double* tsec, double* Mdot, double* lnT);
/* device time [ms] of the last launch of kernel `which` (0 = cell loop, 1 = face loop, 2 = evolve, 3 = the EOS half
 * of the cell loop: the loop runs as an EOS kernel followed by a neutrino-emissivity kernel), measured with CUDA events on the library's stream; launches = kernels launched by the last call */
double dstar_batch_last_kernel_ms(dstar_batch* b, int which);
int dstar_batch_last_launch_count(dstar_batch* b);
/* zones the last cell pass re-evaluated with plain IEEE divisions because an operand of the branch-free form left its
 * safe range: counts[0] = EOS half (or the fused kernel), counts[1] = neutrino half.  Normally zero. */
int dstar_batch_last_redo_counts(dstar_batch* b, int32_t* counts);

/* ---- one-shot host-buffer forms (what the Fortran shim calls for a single star) ---- */
int dstar_micro_tables(dstar_ctx* ctx, int n_star, const int32_t* zone_offset, int ncharged, const double* rho,
                       const double* rho_bar, const double* P, const double* P_bar, const double* dm,
                       const double* ionic, const double* ionic_bar, const double* Yion, const double* Yion_bar,
                       const double* Zion, const double* Aion, const double* Tc_cell, const double* Tc_face,
                       const dstar_micro_ctrl* ctrl, double* tab_lnT, double* tab_lnCp, double* tab_lnGamma,
                       double* tab_lnEnu, double* tab_lnK, double* rho_bar1, int32_t* status);

/* ---- test support: one dsmath function on the device (fn = 0 exp, 1 log, 2 log10, 3 pow(x,y), 4 sin,
 * 5 cos, 6 sinh, 7 cosh, 8 atan; 9-13 exercise the branch-free division and square root of dconst.cuh); dsmath is the deterministic elementary-function layer that plays the
 * role of MESA's math_lib for this implementation (dstar_b200/csrc/dsmath.cuh) ---- */
int dstar_dsmath_eval(dstar_ctx* ctx, int fn, int n, const double* x, const double* y, double* out);
/* test support: send every zone of the cell loop through the plain-division second pass as well (the pass that normally
 * re-evaluates only the zones whose branch-free divisions left their safe operand range) */
void dstar_debug_force_redo(int on);

/* ---- measurement helpers ---- */
/* FP64 DFMA-chain throughput of the device in TFLOP/s (the FP64 roofline denominator) */
int dstar_fp64_peak_tflops(dstar_ctx* ctx, double* tflops);
int dstar_device_info(dstar_ctx* ctx, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char name[256]);

#ifdef __cplusplus
}
#endif
#endif /* DSTAR_B200_H */
