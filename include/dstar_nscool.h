/* dstar_b200 -- host-side mirror of the reference's module API for the hot path.
 *
 * The reference's host language is Fortran (NScool/public/NScool_lib.f); no Fortran compiler exists
 * in the build image, so the host layer above the kernel ABI (dstar_b200.h) is C++ with the same
 * entry points, argument meaning and `ierr` conventions.  A Fortran program binds these through
 * ISO_C_BINDING exactly like dstar_b200.h.
 */
#ifndef DSTAR_NSCOOL_H
#define DSTAR_NSCOOL_H
#include <stdint.h>

#include "dstar_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data files (replace the *_startup readers) ---- */
/* read_helm_table (dStar_eos/private/helm_alloc.f:120): binary written by tools/gen_helm_table */
int dstar_ctx_load_helm_file(dstar_ctx* ctx, const char* path);
/* load_sf_table (superfluid/private/load_sf.f:5): Tc_data text file -> monotone cubic -> device */
int dstar_ctx_load_gap_file(dstar_ctx* ctx, int type, const char* path);
/* MESA interp_pm on (4,n) column-major f with f(1,:) preloaded (host utility) */
int dstar_host_interp_pm(const double* x, int n, double* f);

/* Crust composition from an "abuntime" network-output file (format: dStar_crust/preprocessor/README:3-31):
 * read_abuntime + reduce_abuntime + extend_abuntime of dStar_crust/preprocessor/src/abuntime.f:18-268, driven as
 * process_abuntime.f:66-76 does (points culled to lgP_increment, nuclides below abundance_threshold x the most
 * abundant non-neutron species at every depth dropped, table extended to lgP 22..33.5 and blended into HZ90 over
 * 0.2 dex above the file's last point).  Outputs: network (names: 6 bytes each, NUL-padded; Z, A), lgP(nz),
 * Y(nnet, nz) column-major.  Call with all output buffers NULL to query nnet and nz.  Host only (no device needed).
 * NScool selects such a table with crust_composition = '<name>' -> <data_dir>/crust_data/<name>[.data]. */
int dstar_abuntime_table(const char* path, double lgP_increment, double abundance_threshold, int32_t* nnet, int32_t* nz,
                         char* names, double* Z, double* A, double* lgP, double* Y, int cap_net, int cap_z);

/* ---- NScool_lib (NScool/public/NScool_lib.f:6-91) ---- */
/* NScool_init(my_dStar_dir, ierr): data_dir holds helm_table.bin and Tc_data/ */
int dstar_NScool_init(const char* data_dir, int device);
void dstar_NScool_shutdown(void);
/* alloc_NScool(ierr) -> id >= 1, or < 0 */
int dstar_alloc_NScool(void);
int dstar_dealloc_NScool(int id);
/* NScool_setup(id, inlist, ierr): read the &controls namelist */
int dstar_NScool_setup(int id, const char* inlist);
/* set one control by name after setup (what a driver does through get_NScool_info_ptr) */
int dstar_NScool_set_control(int id, const char* name, const char* value);
/* NScool_create_model(id, ierr) / NScool_evolve_model(id, ierr) for ONE star */
int dstar_NScool_create_model(int id);
int dstar_NScool_evolve_model(int id);
/* the same two calls for MANY stars at once (a sweep / MCMC batch): one coefficient-pass launch pair and
 * one evolve launch for all of them */
int dstar_NScool_create_models(int n, const int32_t* ids);
int dstar_NScool_evolve_models(int n, const int32_t* ids);
/* The sweep driver over the GPUs of ONE node: create_models + evolve_models for n stars on n_gpu devices.  The stars
 * are dealt out in contiguous blocks (block sizes differ by at most one), one host thread per device works on its
 * block, nothing is exchanged between the devices (stars do not interact, NScool_def.f:160).  It replaces N processes
 * each looping NScool_create_model / NScool_evolve_model over their stars (NScool/public/NScool_lib.f:40-90).
 *   devices       (n_gpu) CUDA device indices, or NULL = the device of dstar_NScool_init followed by the node's other
 *                 devices in index order;  the stars of one block must satisfy create_models' batch rules
 *   t/Teff/Qb_monitor  optional (number_epochs, n) arrays, epoch-major: the monitors of all stars, concatenated in the
 *                 order of `ids` (they are also in the star handles: dstar_NScool_get_monitors)
 *   star_status   optional (n): 0, or the most negative isolve return code of the star's epochs
 * Hooks (dstar_NScool_set_hooks) are called from the worker threads.  Returns as evolve_models: 0, or the first
 * device's error (< 0; message in dstar_NScool_last_error). */
int dstar_NScool_run_batch(int n_gpu, const int32_t* devices, int n, const int32_t* ids, double* t_monitor,
                           double* Teff_monitor, double* Qb_monitor, int32_t* star_status);
/* With the control share_identical_tables = .true. on every star of a batch (a dstar_b200 extension; default .false.,
 * i.e. every star tabulates for itself as NScool_create_model.f:340-441 does), create_models evaluates one cell-table
 * set per distinct (structure, composition, T_c) and one face-table set per distinct (cell set, face composition with
 * its impurity parameter, face T_c); every star still reports its own tables, bit-identical to the unshared ones.
 * table_sets: how many distinct cell / face sets the last create_models evaluated. */
int dstar_NScool_table_sets(int32_t* cell_sets, int32_t* face_sets);
/* thread mapping of the evolve kernel for the following evolve calls (dstar_batch_set_evolve_mapping, dstar_b200.h):
 * 0 automatic, 1 CTA per star, 2 warp per star.  Not a reference control: the results do not depend on it. */
int dstar_NScool_set_evolve_mapping(int mapping);
/* hooks (NScool_def.f:148-155), evaluated on the host before upload:
 *   other_set_Qimp(id): may rewrite ionic_bar%Q via dstar_NScool_set_zone_array(id,"Qimp_bar",...)
 *   other_sf_get_results(id,kp,kn,Tc): called per zone for cells and faces */
typedef void (*dstar_set_Qimp_fn)(int id, int* ierr);
typedef void (*dstar_sf_fn)(int id, double kp, double kn, double Tc[3]);
int dstar_NScool_set_hooks(int id, dstar_set_Qimp_fn qimp, dstar_sf_fn sf);
/* native versions of the example hooks alt_Qimp / alt_sf (examples/custom_Tc_Qimp/src/alt_micro.f:4-44),
 * reading extra_real_controls(1:2) = (rho_C, Q_pasta), for sweep drivers */
void dstar_example_alt_Qimp(int id, int* ierr);
void dstar_example_alt_sf(int id, double kp, double kn, double Tc[3]);
/* accessors to type NScool_info (NScool_def.f:33-157) */
int dstar_NScool_get_int(int id, const char* name, int32_t* value);     /* nz, number_epochs, ncharged, model */
int dstar_NScool_get_real(int id, const char* name, double* value);     /* grav, Mtotal, Rtotal, Teff, Lsurf, ... */
/* zone arrays by name (dm, P, rho, rho_bar, ePhi, ..., lnT, ionic(11,nz), ionic_bar, tab_lnK(4*128,nz), ...) */
int dstar_NScool_get_zone_array(int id, const char* name, double* out, int capacity);
int dstar_NScool_set_zone_array(int id, const char* name, const double* in, int count);
/* monitors (NScool_def.f:136-144), length number_epochs */
int dstar_NScool_get_monitors(int id, double* t_monitor, double* Teff_monitor, double* Qb_monitor);
/* isolve statistics iwork(14:20) of the last evolve (NScool_evolve.f:124-130) */
int dstar_NScool_get_counters(int id, int32_t counters[7]);
/* history rows kept from the last evolve (NScool_history.f:84-86); returns the row count */
int dstar_NScool_get_history(int id, int capacity, int32_t* model, double* time_d, double* lg_Mdot,
                             double* lg_Teff, double* lg_Lsurf, double* lg_Lnu, double* lg_Lnuc);
/* write LOGS/history.data in the reference's format (NScool_history.f:25-89; readable by tools/reader.py).
 * Returns 0, 1 = written but truncated at the trace capacity (more rows were produced than kept), < 0 = error
 * (including: no rows kept at all because the batch ran with trace_capacity = 0). */
int dstar_NScool_write_history(int id, const char* path);
/* the terminal log of the last evolve (do_write_terminal, NScool/private/NScool_terminal.f:17-57, driven as
 * evaluate_timestep does, NScool_evolve.f:206-217): model, time [s], dt, lg Lsurf, lg Teff, lg Lnu, lg Lnuc,
 * max lg T and lg P there, min lg T and lg P there.  path NULL or "-" = standard output.  Return as write_history. */
int dstar_NScool_write_terminal(int id, const char* path);
/* the same rows as arrays: (capacity) each; returns the number of rows kept */
int dstar_NScool_get_terminal(int id, int capacity, int32_t* model, double* tsec, double* dt, double* lgTmax,
                              double* lgP_Tmax, double* lgTmin, double* lgP_Tmin);

/* Profiles (do_write_profile, NScool/private/NScool_profile.f:27-106; written when mod(model,
 * write_interval_for_profile) == 0, NScool_evolve.f:227-233).  The device keeps ln T of every zone at those
 * models (control `profile_capacity`: profiles kept per star; -1 = 512 for batches of <= 64 stars, else 0).
 * get_profile: header = time[d], core_mass, core_radius, accretion_rate, heating_lum, cooling_lum, total_mass,
 * total_radius; columns(20, nz) zone-major in the reference's column order (zone, mass, dm, area, gravity, eLambda,
 * ePhi, temperature, luminosity, pressure, density, zbar, abar, Xn, Qimp, Gamma, cp, eps_nu, eps_nuc, K).
 * write_profiles: "<directory>/<base_profile_filename><model>" per profile + the manifest
 * "<directory>/<profile_manifest_filename>", same layout as the reference's files (tools/reader.py reads them). */
int dstar_NScool_profile_count(int id, int32_t* kept, int32_t* produced);
int dstar_NScool_get_profile(int id, int index, int32_t* model, double* header, double* columns);
int dstar_NScool_write_profiles(int id, const char* directory);
/* device timing of the last batch calls [ms]: which = 0 cell loop, 1 face loop, 2 evolve */
double dstar_NScool_last_kernel_ms(int which);
/* message of the last failed dstar_NScool_* call */
const char* dstar_NScool_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DSTAR_NSCOOL_H */
