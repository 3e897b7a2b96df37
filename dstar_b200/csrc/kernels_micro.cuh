// Seam A kernels: the NScool coefficient pass (do_setup_crust_transport,
// NScool/private/NScool_create_model.f:289-444) for a whole batch of stars, plus the batched
// point evaluators behind the generic C-ABI entry points and the device-resident crust
// composition lookup (dStar_crust_get_composition_info, dStar_crust/public/dStar_crust_lib.f:109-147).
//
// Mapping: one CTA per (star, zone), 128 threads = the 128 lnT knots of the zone's tables
// (create_model.f:16-18, :340).  Density, composition, chi, kn and Tc are CTA-uniform, only T
// varies across lanes, so the liquid/solid switch (ion.f:55) and the photo-neutrino temperature
// bands (eval_crust_neutrino.f:209-228) flip at most once along the CTA.  The monotone-cubic
// (interp_pm) coefficients are built in shared memory by the same CTA and written out as the
// (4*128, nz) column-major blocks NScool_def.f:130-134 declares, 32 B per lane, 4 KB per CTA,
// fully coalesced.
#pragma once
#include "conductivity.cuh"
#include "dconst.cuh"
#include "eos.cuh"
#include "neutrino.cuh"

constexpr int DS_NTAB = 128;
#ifndef DS_EOS_COMP_X
#define DS_EOS_COMP_X DS_EOS_ALL   // experiments: components the coefficient kernels evaluate
#endif

// device-resident gap tables (superfluid_def.f:14-21): f = (4,nv) monotone-cubic coefficients
struct DsSfDev {
    int loaded[3], nv[3];
    double kF_min[3], kF_max[3], scale[3];
    const double* kF[3];
    const double* f[3];
};

// k (0-based) with x[k] <= xv < x[k+1], clamped to [0, n-2]
DS_D int ds_locate(const double* __restrict__ x, int n, double xv) {
    if (xv <= x[0]) return 0;
    if (xv >= x[n - 1]) return n - 2;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x[mid] <= xv) lo = mid; else hi = mid;
    }
    return lo;
}
// MESA interp_1d interp_value / interp_value_and_slope on (4,n) coefficients
DS_D double ds_interp_value(const double* __restrict__ x, int n, const double* __restrict__ f, double xv) {
    int k = ds_locate(x, n, xv);
    double dx = xv - x[k];
    const double* c = f + 4 * k;
    return c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
}
// sf_get_results (superfluid/public/superfluid_lib.f:65-101)
DS_D void ds_sf_get_results(const DsSfDev& sf, double kFp, double kFn, double Tc[3]) {
#pragma unroll
    for (int id = 0; id < 3; ++id) {
        double kF = (id == 0) ? kFp : kFn;
        double v = 0.0;
        if (sf.loaded[id] && !(kF < sf.kF_min[id] || kF > sf.kF_max[id]))
            v = fmax(ds_interp_value(sf.kF[id], sf.nv[id], sf.f[id], kF), 0.0) * sf.scale[id];
        Tc[id] = v;
    }
}

struct DsMicroArgs {
    int total_zones, ncharged;
    const int* zone_star;    // (total_zones) star index of each flattened zone
    const int* zone_offset;  // (n_star+1)
    const double *rho, *rho_bar, *P, *P_bar, *dm;  // (total_zones)
    const double *ionic, *ionic_bar;               // (11, total_zones)
    const double *Yion, *Yion_bar;                 // (ncharged, total_zones)
    const double *Zion, *Aion;                     // (ncharged)
    const DsIonZ* ionz;                            // (ncharged) host-evaluated functions of Z (eos.cuh)
    double* aux;                                   // (DS_AUX, total_zones) zone-level inputs prepared by ds_zone_aux_kernel
    const double *Tc_cell, *Tc_face;               // (3, total_zones) or null -> gap tables
    const double *tab_lnT, *tab_T, *tab_lgT;       // (128) host-evaluated knots: lnT, dsm::exp(lnT), dsm::log10(T)
    DsEosCtrl eos;
    DsCondCtrl cond;
    int nu_channels[5];
    int turn_on_shell_Urca;
    double shell_Urca_luminosity_coeff, lgP_shell_Urca;
    double* rho_bar1;  // (n_star) in/out
    int* status;       // (n_star) sticky error word
    double *tab_lnCp, *tab_lnGamma, *tab_lnEnu, *tab_lnK;  // (4*128, total_zones)
    int* redo;        // (total_zones) or null: compact list of the zones whose branch-free divisions left the safe range
    int* redo_count;  // (1) length of that list (zeroed before the branch-free pass)
    int force_redo;  // test hook: flag every zone, so the plain pass recomputes everything
    // table sharing (dstar_batch_set_table_sharing): only the zones of the stars that OWN a table set are evaluated
    // (work_list, n_work; null = every zone), and a star's tables start at slot tab_zone0[star] of the compact table
    // arrays (null = its own zone offset)
    const int* work_list;
    int n_work;
    const int* tab_zone0;
    // block-scheduled EOS half (ds_eos_block_kernel): work items = (zone, 32-knot block), item = 4 * zone + block, in
    // four lists of capacity 4 * total_zones each -- all-solid, all-liquid and mixed blocks, then the redo list
    // the neutrino half and the faces run through the same kernel in natural block order (no lists); their redo
    // lists are blk_redo / blk_redo_count, set per launch
    int* blk_redo;
    int* blk_redo_count;
    int* blk_list;
    int* blk_count;   // (4 + 3 * DS_BLK_NKEY): list lengths, then the sort bins
};

// MESA interp_pm (Steffen 1990) on one zone's 128 knots; i = knot index of this lane, v holds f(1,:) on entry.
// hh[i] = x[i+1] - x[i] is common to all zones.  Called by every thread of the CTA (the barriers are CTA-wide);
// `on` masks groups without a zone.
DS_D void ds_group_interp_pm(bool on, int i, const double* __restrict__ hh, double* v, double* sm,
                             double* sl, double* __restrict__ out) {
    const int n = DS_NTAB;
    if (on && i < n - 1) sm[i] = (v[i + 1] - v[i]) / hh[i];
    __syncthreads();
    double s = 0.0;
    if (on) {
        if (i > 0 && i < n - 1) {
            double p = (sm[i - 1] * hh[i] + sm[i] * hh[i - 1]) / (hh[i - 1] + hh[i]);
            double sg = copysign(1.0, sm[i - 1]) + copysign(1.0, sm[i]);
            s = sg * fmin(fmin(fabs(sm[i - 1]), fabs(sm[i])), 0.5 * fabs(p));
        } else if (i == 0) {
            double p1 = sm[0] * (1.0 + hh[0] / (hh[0] + hh[1])) - sm[1] * hh[0] / (hh[0] + hh[1]);
            if (p1 * sm[0] <= 0.0) s = 0.0;
            else if (fabs(p1) > 2.0 * fabs(sm[0])) s = 2.0 * sm[0];
            else s = p1;
        } else {
            double pn = sm[n - 2] * (1.0 + hh[n - 2] / (hh[n - 2] + hh[n - 3])) -
                        sm[n - 3] * hh[n - 2] / (hh[n - 2] + hh[n - 3]);
            if (pn * sm[n - 2] <= 0.0) s = 0.0;
            else if (fabs(pn) > 2.0 * fabs(sm[n - 2])) s = 2.0 * sm[n - 2];
            else s = pn;
        }
        sl[i] = s;
    }
    __syncthreads();
    if (on) {
        double c2 = 0.0, c3 = 0.0;
        if (i < n - 1) {
            c2 = (3.0 * sm[i] - 2.0 * sl[i] - sl[i + 1]) / hh[i];
            c3 = (sl[i] + sl[i + 1] - 2.0 * sm[i]) / (hh[i] * hh[i]);
        }
        reinterpret_cast<double4*>(out)[i] = make_double4(v[i], s, c2, c3);
    }
    __syncthreads();
}

// Two tables of the same zone at once (lnCp and lnGamma of the EOS half): the same operations per table as
// ds_group_interp_pm, with the three CTA barriers shared.
DS_D void ds_group_interp_pm2(bool on, int i, const double* __restrict__ hh, double* va, double* vb, double* sma, double* smb,
                              double* sla, double* slb, double* __restrict__ outa, double* __restrict__ outb) {
    const int n = DS_NTAB;
    if (on && i < n - 1) { sma[i] = (va[i + 1] - va[i]) / hh[i]; smb[i] = (vb[i + 1] - vb[i]) / hh[i]; }
    __syncthreads();
    double s2[2] = {0.0, 0.0};
    if (on) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double* sm = q ? smb : sma;
            double s = 0.0;
            if (i > 0 && i < n - 1) {
                double p = (sm[i - 1] * hh[i] + sm[i] * hh[i - 1]) / (hh[i - 1] + hh[i]);
                double sg = copysign(1.0, sm[i - 1]) + copysign(1.0, sm[i]);
                s = sg * fmin(fmin(fabs(sm[i - 1]), fabs(sm[i])), 0.5 * fabs(p));
            } else if (i == 0) {
                double p1 = sm[0] * (1.0 + hh[0] / (hh[0] + hh[1])) - sm[1] * hh[0] / (hh[0] + hh[1]);
                if (p1 * sm[0] <= 0.0) s = 0.0;
                else if (fabs(p1) > 2.0 * fabs(sm[0])) s = 2.0 * sm[0];
                else s = p1;
            } else {
                double pn = sm[n - 2] * (1.0 + hh[n - 2] / (hh[n - 2] + hh[n - 3])) -
                            sm[n - 3] * hh[n - 2] / (hh[n - 2] + hh[n - 3]);
                if (pn * sm[n - 2] <= 0.0) s = 0.0;
                else if (fabs(pn) > 2.0 * fabs(sm[n - 2])) s = 2.0 * sm[n - 2];
                else s = pn;
            }
            (q ? slb : sla)[i] = s;
            s2[q] = s;
        }
    }
    __syncthreads();
    if (on) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double *sm = q ? smb : sma, *sl = q ? slb : sla, *v = q ? vb : va;
            double c2 = 0.0, c3 = 0.0;
            if (i < n - 1) {
                c2 = (3.0 * sm[i] - 2.0 * sl[i] - sl[i + 1]) / hh[i];
                c3 = (sl[i] + sl[i + 1] - 2.0 * sm[i]) / (hh[i] * hh[i]);
            }
            reinterpret_cast<double4*>(q ? outb : outa)[i] = make_double4(v[i], s2[q], c2, c3);
        }
    }
    __syncthreads();
}

// Persistent form: ONE CTA per SM, G groups of 128 lanes; group k of CTA b handles zones
// (b*G + k) + round * gridDim.x * G.  The kernel body is ~430 KB of straight-line SASS executed once per
// evaluation, far larger than the instruction caches: with independent 128-thread CTAs the SM held six CTAs
// at six different program counters and the warps starved on instruction fetch (ncu: stalled_no_instruction
// 11 per issue, profiles/r01_notes.md).  Here all G groups start every round together (CTA barrier), so the
// 4*G warps of an SM walk through the code in lockstep and every instruction line is fetched once per SM.
//
// FAST = true evaluates the EOS and the neutrino emissivities with the branch-free divisions of dconst.cuh and
// appends to the list a.redo every zone in which some lane's operands left their safe range; a second launch
// with FAST = false (plain IEEE divisions) then re-evaluates exactly the listed zones (normally none).  Keeping the re-evaluation in a separate
// launch keeps it out of the hot loop's register allocation (a conditional call inside the loop cost 10 %).
// With FAST = false and a.redo == null the kernel is the single plain pass (used for the faces, whose EOS is
// mostly dead code: only Gamma, Theta and mu_e feed the conductivity).
struct DsZoneIn {   // 16 doubles: the prologue fills it as a flat array (11 moments, then the 5 aux values)
    DsComp c;
    double rho, chi, Tc[3];
};
static_assert(sizeof(DsZoneIn) == 16 * sizeof(double), "DsZoneIn is loaded as 16 doubles");
constexpr int DS_PRE = (int)(sizeof(DsZonePre) / sizeof(double)), DS_HZ = (int)(sizeof(DsHelmZone) / sizeof(double));
constexpr int DS_CP = (int)(sizeof(dsk::DsCondPre) / sizeof(double));
constexpr int DS_AUX = 5 + DS_PRE + DS_HZ + DS_CP + 2;   // rho, chi, Tc(3), DsZonePre, DsHelmZone (raw 8-byte words); faces: DsCondPre, lambda_imp, its ierr
static_assert(sizeof(DsHelmZone) % sizeof(double) == 0, "DsHelmZone is copied as 8-byte words");
constexpr int DS_YMAX = 96;   // species kept in shared memory per group (more: read from global memory)
constexpr int DS_IONZ_SMEM = 20;  // species whose Z-only constants (136 B each) are kept in shared memory

// Zone-level inputs that need transcendental functions and table searches (rho of the face pass with the rho_bar(1)
// fix-up, chi, the three critical temperatures): one thread per zone, once per pass, instead of one lane per group
// doing three binary searches through HBM while the other 127 wait at the round's first barrier.
template <bool FACE>
__global__ void ds_zone_aux_kernel(DsMicroArgs a, DsHelmDev helm, DsSfDev sf) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.total_zones) return;
    const int star = a.zone_star[g];
    const int iz = g - a.zone_offset[star];
    const DsComp c = ds_load_comp((FACE ? a.ionic_bar : a.ionic) + (size_t)11 * g);
    const double rho = FACE ? ((iz == 0) ? a.rho_bar1[star] : a.rho_bar[g]) : a.rho[g];
    const double chi = dse::nuclear_volume_fraction(rho, c, DS_DEFAULT_NUCLEAR_RADIUS);
    double Tc[3];
    const double* Tcin = FACE ? a.Tc_face : a.Tc_cell;
    if (Tcin) { Tc[0] = Tcin[3 * (size_t)g]; Tc[1] = Tcin[3 * (size_t)g + 1]; Tc[2] = Tcin[3 * (size_t)g + 2]; }
    else ds_sf_get_results(sf, 0.0, dse::neutron_wavenumber(rho, c, chi), Tc);
    double* o = a.aux + (size_t)DS_AUX * g;
    o[0] = rho; o[1] = chi; o[2] = Tc[0]; o[3] = Tc[1]; o[4] = Tc[2];
    const DsZonePre zp = ds_zone_pre(rho, c, chi);
    for (int q = 0; q < DS_PRE; ++q) o[5 + q] = reinterpret_cast<const double*>(&zp)[q];
    const DsHelmZone hz = ds_helm_zone(helm, rho, c.A / (1.0 - c.Yn), c.Z);   // as ds_eval_crust_eos passes them
    for (int q = 0; q < DS_HZ; ++q) o[5 + DS_PRE + q] = reinterpret_cast<const double*>(&hz)[q];
    if (FACE) {
        const dsk::DsCondPre cp = dsk::ds_cond_pre(a.cond, rho, c);
        for (int q = 0; q < DS_CP; ++q) o[5 + DS_PRE + DS_HZ + q] = reinterpret_cast<const double*>(&cp)[q];
        // impurity Coulomb logarithm of the neutron channel: T-independent, once per zone (neutron_conductivity.f:382)
        double lam = nan("");
        int lam_ierr = 0;
        if (a.cond.include_neutrons && c.Yn > 0.0) lam_ierr = dsk::ds_lambda_imp(rho / dsc::amu * c.Yn / (1.0 - chi), c, lam);
        o[5 + DS_PRE + DS_HZ + DS_CP] = lam;
        o[5 + DS_PRE + DS_HZ + DS_CP + 1] = (double)lam_ierr;
    }
}

// PART selects what a launch evaluates.  The cell pass of the reference computes the EOS (-> lnCp, lnGamma) and the
// neutrino emissivity (-> lnEnu) per (zone, knot) in one loop body (create_model.f:344-403); the emissivity takes
// nothing from the EOS, so the two halves are independent kernels here (DS_PART_EOS + DS_PART_NU), each with the
// register allocation and group count its own live state wants.  DS_PART_CELL is the fused form (kept selectable:
// DSTAR_MICRO_SPLIT=0), DS_PART_FACE the face pass (EOS -> Gamma, Theta, mu_e -> conductivity -> lnK).
enum { DS_PART_CELL = 0, DS_PART_EOS = 1, DS_PART_NU = 2, DS_PART_FACE = 3 };
__host__ __device__ constexpr int ds_micro_nb(int part) { return part == DS_PART_EOS ? 2 : 1; }   // tables interpolated together (EOS half: lnCp, lnGamma)
// dynamic shared memory of ds_micro_kernel<PART, G, *>
__host__ __device__ constexpr size_t ds_micro_dyn_bytes(int part, int G) {
    return sizeof(double) * ((size_t)2 * G * ds_micro_nb(part) * 128 + (part == DS_PART_NU ? 0 : (size_t)G * 96));
}

#ifndef DS_MICRO_CTAS
#define DS_MICRO_CTAS 1   // experiments: persistent CTAs per SM
#endif
template <int PART, int G, bool FAST>
__global__ void __launch_bounds__(DS_NTAB* G, DS_MICRO_CTAS) ds_micro_kernel(DsMicroArgs a, DsHelmDev helm, DsSfDev sf) {
    constexpr bool FACE = PART == DS_PART_FACE;
    constexpr bool HAS_EOS = PART != DS_PART_NU;                          // evaluates eval_crust_eos
    constexpr bool HAS_NU = PART == DS_PART_CELL || PART == DS_PART_NU;   // evaluates the crust neutrino emissivity
    constexpr bool SPECIES = PART == DS_PART_CELL || PART == DS_PART_EOS; // enters the species loop of ion_mixture
    constexpr int NV = PART == DS_PART_CELL ? 3 : PART == DS_PART_EOS ? 2 : 1;
    constexpr int V_ENU = PART == DS_PART_CELL ? 2 : 0;   // row of s_v holding lnEnu
    constexpr int NB = ds_micro_nb(PART);
    __shared__ double s_h[DS_NTAB], s_v[G][NV][DS_NTAB];
    __shared__ double s_lam[G];
    __shared__ int s_lam_ierr[G], s_err[G], s_bad[G];
    __shared__ DsZoneIn s_zone[G];
    // dynamic part (the static 48 KB are used up): interpolation scratch, then G * DS_YMAX abundances (none for DS_PART_NU)
    extern __shared__ double s_dyn[];
    double (*s_m)[NB][DS_NTAB] = reinterpret_cast<double (*)[NB][DS_NTAB]>(s_dyn);
    double (*s_s)[NB][DS_NTAB] = reinterpret_cast<double (*)[NB][DS_NTAB]>(s_dyn + G * NB * DS_NTAB);
    double (*s_Y)[DS_YMAX] = reinterpret_cast<double (*)[DS_YMAX]>(s_dyn + 2 * G * NB * DS_NTAB);
    __shared__ unsigned char s_act[SPECIES ? G : 1][DS_YMAX];   // indices of the zone's species above the abundance threshold
    __shared__ int s_nact[G];
    __shared__ DsZonePre s_pre[G];
    __shared__ DsHelmZone s_hz[HAS_EOS ? G : 1];
    __shared__ dsk::DsCondPre s_cp[FACE ? G : 1];
    __shared__ double s_ZA[2][HAS_EOS ? DS_YMAX : 1];       // charge and mass numbers of the network
    __shared__ DsIonZ s_ionz[HAS_EOS ? DS_IONZ_SMEM : 1];   // per-species constants of the first species (the rest: global)
    const int grp = threadIdx.x / DS_NTAB, it = threadIdx.x % DS_NTAB;
    if (threadIdx.x < DS_NTAB - 1) s_h[it] = a.tab_lnT[it + 1] - a.tab_lnT[it];
    const bool net_in_smem = HAS_EOS && a.ncharged <= DS_YMAX, ionz_in_smem = HAS_EOS && a.ncharged <= DS_IONZ_SMEM;
    if (net_in_smem)
        for (int i = threadIdx.x; i < a.ncharged; i += blockDim.x) { s_ZA[0][i] = a.Zion[i]; s_ZA[1][i] = a.Aion[i]; }
    if (ionz_in_smem)
        for (int i = threadIdx.x; i < a.ncharged * (int)(sizeof(DsIonZ) / sizeof(double)); i += blockDim.x)
            reinterpret_cast<double*>(s_ionz)[i] = reinterpret_cast<const double*>(a.ionz)[i];
    const double* Zn = net_in_smem ? s_ZA[0] : a.Zion;
    const double* An = net_in_smem ? s_ZA[1] : a.Aion;
    const DsIonZ* ionz = ionz_in_smem ? s_ionz : a.ionz;
    // Which 32-knot block a warp of the group works on is rotated by the group index: the warp whose knots straddle
    // the melting line (it runs both ion branches) sits at a similar temperature in most zones, i.e. would be the
    // same warp number in every group -- and warp number mod 4 is the SM sub-partition.  The rotation spreads those
    // double-work warps over the four sub-partitions.  Knots stay contiguous within a warp.
    const int kn = ((((it >> 5) + grp) & 3) << 5) | (it & 31);
    const double T = a.tab_T[kn], lgT = a.tab_lgT[kn];
    // second pass (FAST = false with a redo list): the work items are the list entries, densely packed
    const bool redo_pass = !FAST && a.redo != nullptr;
    const int n_items = redo_pass ? *a.redo_count : (a.work_list ? a.n_work : a.total_zones);
    for (int g0 = blockIdx.x * G; g0 < n_items; g0 += gridDim.x * G) {
        const bool on = g0 + grp < n_items;
        const int gq = redo_pass ? (on ? a.redo[g0 + grp] : 0) : (a.work_list ? (on ? a.work_list[g0 + grp] : 0) : g0 + grp);
        if (it == 0) { s_err[grp] = 0; s_lam[grp] = nan(""); s_lam_ierr[grp] = 0; s_bad[grp] = 0; }
        __syncthreads();   // start of round: all groups aligned
        // groups past the end re-evaluate the last zone (results discarded) so that every thread of the
        // CTA reaches the DS_PHASE alignment points inside the evaluators
        const int g = on ? gq : a.total_zones - 1;
        // zone-level inputs live ONCE per group in shared memory (all 128 lanes read them by broadcast) instead of
        // 16 doubles in every lane's registers: the evaluators are register-starved at 72 registers per thread
        const int star = a.zone_star[g];
        const int iz = g - a.zone_offset[star];
        const int nz = a.zone_offset[star + 1] - a.zone_offset[star];
        // prologue: 16 lanes fetch the zone's 11 composition moments and 5 prepared values, the next lanes its
        // abundances -- independent loads, one memory round trip per round
        if (it < 11) reinterpret_cast<double*>(&s_zone[grp])[it] = ((FACE ? a.ionic_bar : a.ionic) + (size_t)11 * g)[it];
        else if (it < 16) reinterpret_cast<double*>(&s_zone[grp])[it] = a.aux[(size_t)DS_AUX * g + (it - 11)];
        else if (it < 16 + DS_PRE) reinterpret_cast<double*>(&s_pre[grp])[it - 16] = a.aux[(size_t)DS_AUX * g + (it - 11)];
        else if (HAS_EOS && it < 16 + DS_PRE + DS_HZ)
            reinterpret_cast<double*>(&s_hz[HAS_EOS ? grp : 0])[it - 16 - DS_PRE] = a.aux[(size_t)DS_AUX * g + (it - 11)];
        else if (FACE && it < 16 + DS_PRE + DS_HZ + DS_CP)
            reinterpret_cast<double*>(&s_cp[FACE ? grp : 0])[it - 16 - DS_PRE - DS_HZ] = a.aux[(size_t)DS_AUX * g + (it - 11)];
        const double* Yg = (FACE ? a.Yion_bar : a.Yion) + (size_t)a.ncharged * g;
        if (HAS_EOS && a.ncharged <= DS_YMAX && it < a.ncharged) s_Y[grp][it] = Yg[it];
        __syncthreads();
        if (SPECIES) {   // (the face pass never enters the species loop: its ion thermodynamics are dead code)
            if (a.ncharged <= DS_YMAX && it == 0) {   // active-species list of the zone (ascending, as the reference loops)
                int na = 0;
                for (int i = 0; i < a.ncharged; ++i) if (s_Y[grp][i] > a.eos.Ythresh) s_act[SPECIES ? grp : 0][na++] = (unsigned char)i;
                s_nact[grp] = na;
            }
            __syncthreads();
        }
        const DsComp& c = s_zone[grp].c;
        const DsZonePre& zp = s_pre[grp];
        const double& rho = s_zone[grp].rho;
        const double& chi = s_zone[grp].chi;
        const double* Tc = s_zone[grp].Tc;
        if (FACE) {
            if (a.cond.include_neutrons && c.Yn > 0.0 && it == 0) {
                // impurity Coulomb logarithm: T-independent, once per zone (neutron_conductivity.f:382)
                double nn = rho / dsc::amu * c.Yn / (1.0 - chi);
                double lam;
                s_lam_ierr[grp] = dsk::ds_lambda_imp(nn, c, lam);
                s_lam[grp] = lam;
            }
            __syncthreads();
        }
        bool bad = false;  // branch-free divisions left their safe operand range somewhere (dconst.cuh)
        if (HAS_EOS) {
            // the 16 density weights of the HELM interpolation are read in place (shared memory, broadcast): a register
            // copy per lane added 150 bytes to the spill frame (EOS half 4.39 -> 4.12 ms without it)
            const DsHelmZone& hz = s_hz[HAS_EOS ? grp : 0];
            DsEosRes r;
            const double* Yz = (a.ncharged <= DS_YMAX) ? s_Y[grp] : Yg;
            ds_eval_crust_eos<FAST, DS_EOS_COMP_X>(helm, a.eos, rho, T, lgT, c, zp, hz, a.ncharged, Zn, An, ionz, Yz, Tc[1], chi, r, bad, nullptr,
                                    (SPECIES && a.ncharged <= DS_YMAX) ? s_act[SPECIES ? grp : 0] : nullptr, SPECIES ? s_nact[grp] : 0);
            if (on && r.ierr != 0) s_err[grp] = r.ierr;
            if (!FACE) {
                s_v[grp][0][kn] = dsm::log(r.Cp);
                s_v[grp][NV > 1 ? 1 : 0][kn] = dsm::log(r.Gamma);
                if (on && iz == 0 && kn == 0)  // create_model.f:367-369
                    a.rho_bar1[star] = rho * dsm::pow(a.P_bar[g] / a.P[g], 1.0 / r.chiRho);
            } else {
                DsCond K;
                dsk::ds_conductivity(a.cond, rho, T, chi, r.Gamma, r.Theta, r.mu_e, c, s_cp[FACE ? grp : 0], Tc[1], s_lam[grp],
                                     s_lam_ierr[grp], K);
                s_v[grp][0][kn] = dsm::log(K.total);
            }
        }
        if (HAS_NU) {
            DsNeutrino eps;
            dsn::ds_crust_neutrino<FAST>(rho, T, c, zp, chi, Tc[1], a.nu_channels, eps, bad);
            double lnenu = dsm::log(eps.total / rho);
            if (a.turn_on_shell_Urca && iz < nz - 1) {  // create_model.f:379-384
                if (dsm::log10(a.P_bar[g]) < a.lgP_shell_Urca && dsm::log10(a.P_bar[g + 1]) > a.lgP_shell_Urca)
                    lnenu = dsm::log(dsm::exp(lnenu) + 1.0e36 * a.shell_Urca_luminosity_coeff *
                                                 dsd::ipow<5>(T / 5.0e8) / a.dm[g]);
            }
            s_v[grp][V_ENU][kn] = lnenu;
        }
        if (FAST && bad) s_bad[grp] = 1;
        __syncthreads();
        const size_t ob = (size_t)4 * DS_NTAB * (a.tab_zone0 ? a.tab_zone0[star] + iz : g);
        if (HAS_NU) ds_group_interp_pm(on, kn, s_h, s_v[grp][V_ENU], s_m[grp][0], s_s[grp][0], a.tab_lnEnu + ob);
        if (PART == DS_PART_EOS) {
            ds_group_interp_pm2(on, kn, s_h, s_v[grp][0], s_v[grp][NV > 1 ? 1 : 0], s_m[grp][0], s_m[grp][NB - 1], s_s[grp][0],
                                s_s[grp][NB - 1], a.tab_lnCp + ob, a.tab_lnGamma + ob);
        } else if (HAS_EOS && !FACE) {
            ds_group_interp_pm(on, kn, s_h, s_v[grp][0], s_m[grp][0], s_s[grp][0], a.tab_lnCp + ob);
            ds_group_interp_pm(on, kn, s_h, s_v[grp][NV > 1 ? 1 : 0], s_m[grp][0], s_s[grp][0], a.tab_lnGamma + ob);
        }
        if (FACE) ds_group_interp_pm(on, kn, s_h, s_v[grp][0], s_m[grp][0], s_s[grp][0], a.tab_lnK + ob);
        if (on && it == 0 && s_err[grp] != 0) a.status[star] = s_err[grp];
        if (FAST && on && it == 0 && (s_bad[grp] | a.force_redo)) a.redo[atomicAdd(a.redo_count, 1)] = g;
    }
}

// ---------------------------------------------------------------- EOS half, scheduled by 32-knot blocks
// In ds_micro_kernel a zone is a group of four warps, and its interp_pm needs all four: the warp whose 32 knots
// straddle the melting line runs the liquid AND the solid ion fits, the other three wait for it at the next barrier,
// every round (9 % of the EOS kernel's stall samples at that one barrier, profiles/r02_ncu_full_summary.md).  Here the
// evaluation and the interpolation are separate kernels, so the unit of work is one warp = one (zone, 32-knot block);
// the blocks are sorted into all-solid, all-liquid and mixed ones (ds_eos_classify_kernel) and a persistent CTA takes
// DS_BLK_W blocks of ONE class per round: every warp of a round has the same work.  The class only steers the schedule: a
// misclassified block is evaluated correctly, its warp just diverges.
// EOS half, warps per CTA 16 / 20 / 24 / 28 (128 / 96 / 80 / 72 registers): 3.36 / 3.38 / 3.41 / 3.47 ms; INT-16 (80 species)
// 10.34 / 10.10 / 10.10 / - ms: 20.  The neutrino and face units set their own (32 and 28).
#ifndef DS_BLK_W_X
#define DS_BLK_W_X 20
#endif
constexpr int DS_BLK_W = DS_BLK_W_X;   // warps (blocks) per CTA and round

// Sort key of a block: its phase class (0 all solid, 1 all liquid, 2 mixed -- the phase of every knot as ion_mixture
// decides it, ion.f:55 with lattice_properties' Gamma_e) and the number of species above the abundance threshold
// (the trip count of ion_mixture's loop), capped.  Blocks are laid out class by class and, inside a class, by
// ascending species count, so that the 28 blocks of a round cost the same.
constexpr int DS_BLK_NKEY = 64;   // species-count keys per class
struct DsBlkKey { int cls[4], nkey; };
DS_D DsBlkKey ds_blk_keys(const DsMicroArgs& a, int g) {
    DsBlkKey k;
    const double Z53 = a.ionic[(size_t)11 * g + 2];
    const double rs = a.aux[(size_t)DS_AUX * g + 5 + 1];   // DsZonePre::rs
    // liquid iff Gamma_e Z53 < Gamma_melt, Gamma_e = 2 Ry / (kB T rs) (lattice_properties, dStar_eos_lib.f:84-98), i.e.
    // iff T > T_melt.  The class only steers the schedule, so the comparison need not round as ion_mixture's does.
    const double T_melt = 2.0 * dsc::Rydberg / dsc::boltzmann / rs * Z53 / a.eos.Gamma_melt;
    for (int b = 0; b < 4; ++b) {
        int n_solid = 0;
        for (int q = 0; q < 32; ++q) n_solid += !(a.tab_T[32 * b + q] > T_melt);
        k.cls[b] = (n_solid == 32) ? 0 : (n_solid == 0) ? 1 : 2;
    }
    int na = 0;
    const double* Yg = a.Yion + (size_t)a.ncharged * g;
    for (int i = 0; i < a.ncharged; ++i) na += Yg[i] > a.eos.Ythresh;
    k.nkey = min(na, DS_BLK_NKEY - 1);
    return k;
}
// Counting sort in three launches: histogram of the keys (MODE 0), exclusive prefix per class (ds_eos_prefix_kernel),
// scatter (MODE 1).  blk_count: [0..3] list lengths (3 = redo list), then DS_BLK_NKEY bins per class.  One atomic per
// warp and distinct key (the bins of a real crust are few and hot).
template <int MODE>
__global__ void ds_eos_classify_kernel(DsMicroArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_items = a.work_list ? a.n_work : a.total_zones;
    const bool live = i < n_items;
    const int g = live ? (a.work_list ? a.work_list[i] : i) : 0;
    const DsBlkKey k = ds_blk_keys(a, g);
    const int lane = threadIdx.x & 31;
    for (int b = 0; b < 4; ++b) {
        const int bin = live ? k.cls[b] * DS_BLK_NKEY + k.nkey : -1;
        const unsigned m = __match_any_sync(0xffffffffu, bin);
        if (!live) continue;
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(a.blk_count + 4 + bin, __popc(m));
        base = __shfl_sync(m, base, leader);
        if (MODE == 1)
            a.blk_list[(size_t)k.cls[b] * 4 * a.total_zones + base + __popc(m & ((1u << lane) - 1u))] = 4 * g + b;
    }
}
// bins -> start offsets (in place) and list lengths: one thread per class
template <int UNIT = 0>
__global__ void ds_eos_prefix_kernel(int* blk_count) {
    const int c = threadIdx.x;
    if (c >= 3) return;
    int run = 0;
    for (int q = 0; q < DS_BLK_NKEY; ++q) { const int n = blk_count[4 + c * DS_BLK_NKEY + q]; blk_count[4 + c * DS_BLK_NKEY + q] = run; run += n; }
    blk_count[c] = run;
}

template <int PART, bool FAST>
__global__ void __launch_bounds__(32 * DS_BLK_W, 1) ds_block_kernel(DsMicroArgs a, DsHelmDev helm) {
    constexpr int W = DS_BLK_W;
    constexpr bool FACE = PART == DS_PART_FACE, HAS_EOS = PART != DS_PART_NU, SPECIES = PART == DS_PART_EOS;
    __shared__ DsZoneIn s_zone[W];
    __shared__ DsZonePre s_pre[W];
    __shared__ DsHelmZone s_hz[HAS_EOS ? W : 1];
    __shared__ dsk::DsCondPre s_cp[FACE ? W : 1];
    __shared__ double s_lam[FACE ? W : 1][2];
    __shared__ unsigned char s_act[SPECIES ? W : 1][DS_YMAX];
    __shared__ int s_nact[W];
    __shared__ double s_ZA[2][SPECIES ? DS_YMAX : 1];
    __shared__ DsIonZ s_ionz[SPECIES ? DS_IONZ_SMEM : 1];
    extern __shared__ double s_dyn[];   // EOS half: W * DS_YMAX abundances
    double (*s_Y)[DS_YMAX] = reinterpret_cast<double (*)[DS_YMAX]>(s_dyn);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool net_in_smem = SPECIES && a.ncharged <= DS_YMAX, ionz_in_smem = SPECIES && a.ncharged <= DS_IONZ_SMEM;
    if (net_in_smem)
        for (int i = threadIdx.x; i < a.ncharged; i += blockDim.x) { s_ZA[0][i] = a.Zion[i]; s_ZA[1][i] = a.Aion[i]; }
    if (ionz_in_smem)
        for (int i = threadIdx.x; i < a.ncharged * (int)(sizeof(DsIonZ) / sizeof(double)); i += blockDim.x)
            reinterpret_cast<double*>(s_ionz)[i] = reinterpret_cast<const double*>(a.ionz)[i];
    __syncthreads();
    const double* Zn = net_in_smem ? s_ZA[0] : a.Zion;
    const double* An = net_in_smem ? s_ZA[1] : a.Aion;
    const DsIonZ* ionz = ionz_in_smem ? s_ionz : a.ionz;
    const int n_zones = a.work_list ? a.n_work : a.total_zones;
    // EOS half, branch-free pass: the three class lists one after the other.  Otherwise one list: the redo list
    // (plain pass after a branch-free one) or the natural order, 4 blocks per work zone (neutrino half, faces).
    const bool redo_pass = !FAST && !FACE;
    const int first = (SPECIES && FAST) ? 0 : 3, last = (SPECIES && FAST) ? 3 : 4;
    for (int li = first; li < last; ++li) {
        const int* list = (SPECIES && FAST) ? a.blk_list + (size_t)li * 4 * a.total_zones : (redo_pass ? a.blk_redo : nullptr);
        const int count = (SPECIES && FAST) ? a.blk_count[li] : (redo_pass ? *a.blk_redo_count : 4 * n_zones);
        for (int base = blockIdx.x * W; base < count; base += gridDim.x * W) {
            // With the neutron / phonon channels on, a face evaluation is ~25 k instructions per lane (two DOP853
            // quadratures): free-running warps then thrash the instruction caches (accretion faces 29.7 ms instead of
            // 13.3), so the warps of a CTA start such rounds together.
            if (FACE && (a.cond.include_neutrons | a.cond.include_sf_phonons)) __syncthreads();
            // warps past the end of the list redo its last block and discard it (the EOS half's evaluators hold
            // CTA-wide alignment barriers)
            const bool on = base + w < count;
            const int idx = on ? base + w : count - 1;
            const int item = list ? list[idx] : 4 * (a.work_list ? a.work_list[idx >> 2] : (idx >> 2)) + (idx & 3);
            const int g = item >> 2, kn = ((item & 3) << 5) | lane;
            const int star = a.zone_star[g];
            const int iz = g - a.zone_offset[star];
            const int nz = a.zone_offset[star + 1] - a.zone_offset[star];
            // the zone's inputs into this warp's slots (the slots are the warp's own: no CTA barrier needed)
            const double* ion_g = (FACE ? a.ionic_bar : a.ionic) + (size_t)11 * g;
            constexpr int NIN = 16 + DS_PRE + (HAS_EOS ? DS_HZ : 0) + (FACE ? DS_CP + 2 : 0);
            for (int q = lane; q < NIN; q += 32) {
                const double v = (q < 11) ? ion_g[q] : a.aux[(size_t)DS_AUX * g + (q - 11)];
                if (q < 16) reinterpret_cast<double*>(&s_zone[w])[q] = v;
                else if (q < 16 + DS_PRE) reinterpret_cast<double*>(&s_pre[w])[q - 16] = v;
                else if (q < 16 + DS_PRE + DS_HZ) reinterpret_cast<double*>(&s_hz[HAS_EOS ? w : 0])[q - 16 - DS_PRE] = v;
                else if (q < 16 + DS_PRE + DS_HZ + DS_CP) reinterpret_cast<double*>(&s_cp[FACE ? w : 0])[q - 16 - DS_PRE - DS_HZ] = v;
                else s_lam[FACE ? w : 0][q - 16 - DS_PRE - DS_HZ - DS_CP] = v;
            }
            const double* Yg = (FACE ? a.Yion_bar : a.Yion) + (size_t)a.ncharged * g;
            if (net_in_smem) {   // abundances and the list of the species above the threshold (ascending, as the reference loops)
                int na = 0;
                for (int i0 = 0; i0 < a.ncharged; i0 += 32) {
                    const int i = i0 + lane;
                    const double y = (i < a.ncharged) ? Yg[i] : 0.0;
                    if (i < a.ncharged) s_Y[w][i] = y;
                    const unsigned m = __ballot_sync(0xffffffffu, i < a.ncharged && y > a.eos.Ythresh);
                    if ((m >> lane) & 1u) s_act[SPECIES ? w : 0][na + __popc(m & ((1u << lane) - 1u))] = (unsigned char)i;
                    na += __popc(m);
                }
                if (lane == 0) s_nact[w] = na;
            }
            __syncwarp();
            const DsComp& c = s_zone[w].c;
            const double rho = s_zone[w].rho, chi = s_zone[w].chi;
            const double T = a.tab_T[kn], lgT = a.tab_lgT[kn];
            bool bad = false;
            double v0 = 0.0, v1 = 0.0;
            int ierr = 0;
            if (HAS_EOS) {
                DsEosRes r;
                ds_eval_crust_eos<FAST>(helm, a.eos, rho, T, lgT, c, s_pre[w], s_hz[HAS_EOS ? w : 0], a.ncharged, Zn, An, ionz,
                                        net_in_smem ? s_Y[w] : Yg, s_zone[w].Tc[1], chi, r, bad, nullptr,
                                        net_in_smem ? s_act[SPECIES ? w : 0] : nullptr, net_in_smem ? s_nact[w] : 0);
                ierr = r.ierr;
                if (!FACE) {
                    v0 = dsm::log(r.Cp); v1 = dsm::log(r.Gamma);
                    if (on && iz == 0 && kn == 0 && !(FAST && (bad || a.force_redo)))  // create_model.f:367-369
                        a.rho_bar1[star] = rho * dsm::pow(a.P_bar[g] / a.P[g], 1.0 / r.chiRho);
                } else {
                    DsCond K;
                    dsk::ds_conductivity(a.cond, rho, T, chi, r.Gamma, r.Theta, r.mu_e, c, s_cp[FACE ? w : 0], s_zone[w].Tc[1],
                                         s_lam[FACE ? w : 0][0], (int)s_lam[FACE ? w : 0][1], K);
                    v0 = dsm::log(K.total);
                }
            } else {
                DsNeutrino eps;
                dsn::ds_crust_neutrino<FAST>(rho, T, c, s_pre[w], chi, s_zone[w].Tc[1], a.nu_channels, eps, bad);
                double lnenu = dsm::log(eps.total / rho);
                if (a.turn_on_shell_Urca && iz < nz - 1) {  // create_model.f:379-384
                    if (dsm::log10(a.P_bar[g]) < a.lgP_shell_Urca && dsm::log10(a.P_bar[g + 1]) > a.lgP_shell_Urca)
                        lnenu = dsm::log(dsm::exp(lnenu) + 1.0e36 * a.shell_Urca_luminosity_coeff *
                                                     dsd::ipow<5>(T / 5.0e8) / a.dm[g]);
                }
                v0 = lnenu;
            }
            // the table values of this knot are parked at the head of the zone's (4,128) blocks (contiguous in the
            // knot: coalesced here and in the interpolation kernels, which complete the blocks)
            const size_t ob = (size_t)4 * DS_NTAB * (a.tab_zone0 ? a.tab_zone0[star] + iz : g);
            const bool flagged = FAST && (__any_sync(0xffffffffu, bad) || a.force_redo);
            if (on && !flagged) {
                if (PART == DS_PART_EOS) { a.tab_lnCp[ob + kn] = v0; a.tab_lnGamma[ob + kn] = v1; }
                else if (PART == DS_PART_NU) a.tab_lnEnu[ob + kn] = v0;
                else a.tab_lnK[ob + kn] = v0;
            }
            if (on && ierr != 0) a.status[star] = ierr;
            if (on && flagged && lane == 0) a.blk_redo[atomicAdd(a.blk_redo_count, 1)] = item;
            __syncwarp();   // the slots are rewritten in the next round
        }
    }
}

// interp_pm of lnCp and lnGamma for the zones of the work list: DS_INTERP_G zones per CTA, 128 threads each.  The values
// are read from the head of each block (where ds_eos_block_kernel parked them) before any of it is overwritten.
// zones per CTA: 1 / 2 / 4 -> the three interpolation launches of a step together 0.40 / 0.42 / 0.47 ms (the phases of many
// small CTAs overlap better)
#ifndef DS_INTERP_G_X
#define DS_INTERP_G_X 1
#endif
constexpr int DS_INTERP_G = DS_INTERP_G_X;
template <int UNIT = 0>
__global__ void __launch_bounds__(DS_NTAB* DS_INTERP_G) ds_eos_interp_kernel(DsMicroArgs a) {
    constexpr int G = DS_INTERP_G;
    __shared__ double s_h[DS_NTAB], s_v[G][2][DS_NTAB], s_m[G][2][DS_NTAB], s_s[G][2][DS_NTAB];
    const int grp = threadIdx.x / DS_NTAB, i = threadIdx.x % DS_NTAB;
    const int n_items = a.work_list ? a.n_work : a.total_zones;
    const int idx = blockIdx.x * G + grp;
    const bool on = idx < n_items;
    const int g = on ? (a.work_list ? a.work_list[idx] : idx) : 0;
    const int star = a.zone_star[g];
    const int iz = g - a.zone_offset[star];
    const size_t ob = (size_t)4 * DS_NTAB * (a.tab_zone0 ? a.tab_zone0[star] + iz : g);
    if (threadIdx.x < DS_NTAB - 1) s_h[i] = a.tab_lnT[i + 1] - a.tab_lnT[i];
    s_v[grp][0][i] = on ? a.tab_lnCp[ob + i] : 0.0;
    s_v[grp][1][i] = on ? a.tab_lnGamma[ob + i] : 0.0;
    __syncthreads();
    ds_group_interp_pm2(on, i, s_h, s_v[grp][0], s_v[grp][1], s_m[grp][0], s_m[grp][1], s_s[grp][0], s_s[grp][1],
                        a.tab_lnCp + ob, a.tab_lnGamma + ob);
}

// interp_pm of ONE table (lnEnu or lnK) whose values the block kernel parked at the head of each zone's block
template <int UNIT = 0>
__global__ void __launch_bounds__(DS_NTAB* DS_INTERP_G) ds_interp1_kernel(DsMicroArgs a, double* __restrict__ tab) {
    constexpr int G = DS_INTERP_G;
    __shared__ double s_h[DS_NTAB], s_v[G][DS_NTAB], s_m[G][DS_NTAB], s_s[G][DS_NTAB];
    const int grp = threadIdx.x / DS_NTAB, i = threadIdx.x % DS_NTAB;
    const int n_items = a.work_list ? a.n_work : a.total_zones;
    const int idx = blockIdx.x * G + grp;
    const bool on = idx < n_items;
    const int g = on ? (a.work_list ? a.work_list[idx] : idx) : 0;
    const int star = a.zone_star[g];
    const int iz = g - a.zone_offset[star];
    const size_t ob = (size_t)4 * DS_NTAB * (a.tab_zone0 ? a.tab_zone0[star] + iz : g);
    if (threadIdx.x < DS_NTAB - 1) s_h[i] = a.tab_lnT[i + 1] - a.tab_lnT[i];
    s_v[grp][i] = on ? tab[ob + i] : 0.0;
    __syncthreads();
    ds_group_interp_pm(on, i, s_h, s_v[grp], s_m[grp], s_s[grp], tab + ob);
}

