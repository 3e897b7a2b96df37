// dstar_b200 C ABI (include/dstar_b200.h): contexts, HBM-resident batches, kernel launches.
// Single translation unit: all device code is header-inlined into the kernels below.
#include <cuda_runtime.h>

#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dstar_b200.h"
#include "atm_launch.hpp"
#include "kernels_evolve.cuh"
#include "kernels_evolve_warp.cuh"
#include "kernels_micro.cuh"
#include "kernels_point.cuh"
#include "kernels_test.cuh"
#include "micro_launch.hpp"

constexpr int DS_N_EVOLVE_CLASS = 6;
constexpr int DS_EVOLVE_Z[DS_N_EVOLVE_CLASS] = {2, 4, 6, 8, 9, 10};   // zones per lane: nz <= 64, 128, 192, 256, 288, 320

namespace {
thread_local std::string g_err;
int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) std::snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else std::snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}
#define CK(call)                                                                 \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) return fail(DSTAR_ERR_CUDA, #call, e_);           \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) {
        if (p) { cudaFree(p); p = nullptr; }
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    cudaError_t ensure(size_t count) { return (p && count <= n) ? cudaSuccess : alloc(count); }   // no realloc per step
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        if (count > n || !p) { cudaError_t e = alloc(count); if (e != cudaSuccess) return e; }
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    cudaError_t download(T* h, size_t count, cudaStream_t s) const {
        return cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s);
    }
};
}  // namespace

struct dstar_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // uploads that only the face pass / the evolve kernel need (overlap the cell pass)
    cudaDeviceProp prop{};
    bool helm_loaded = false;
    DsHelmDev helm{};
    DevBuf<double> helm_node, helm_t, helm_d;
    DsSfDev sf{};
    DevBuf<double> sf_kF[3], sf_f[3];
    DevBuf<double> tab_lnT, tab_T, tab_lgT;
    double h_tab_lnT[DS_NTAB];
    DsHostConsts hc{};   // host-evaluated constants (each translation unit with kernels sets its own __constant__ copy)
};

struct dstar_batch {
    dstar_ctx* ctx = nullptr;
    int n_star = 0, total = 0, ncharged = 0, n_epoch = 0, n_atm = 0, atm_nv = 0, trace_cap = 0;
    int prof_interval = 0, prof_cap = 0;
    DevBuf<int> prof_count, prof_model;
    DevBuf<double> prof_tsec, prof_Mdot, prof_lnT;
    std::vector<int> h_zone_offset;
    int max_nz = 0;
    bool have_structure = false, have_tables = false, have_evolve = false, have_Tc_cell = false, have_Tc_face = false,
         have_enuc = false;
    DevBuf<int> blk_list, blk_count;   // block-scheduled EOS half: four item lists and their lengths
    DevBuf<int> zone_offset, zone_star, status, redo, redo_count, atm_index, idid, counters, model, trace_count, trace_model;
    DevBuf<double> rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, Yion, Yion_bar, Zion, Aion, Tc_cell, Tc_face;
    DevBuf<DsIonZ> ionz;
    std::vector<DsIonZ> h_ionz;
    std::vector<double> h_rb1;
    DevBuf<double> aux;   // (DS_AUX, total) zone-level inputs of the coefficient pass (ds_zone_aux_kernel)
    DevBuf<double> rho_bar1, tab_lnCp, tab_lnGamma, tab_lnEnu, tab_lnK;
    DevBuf<double> dm_bar, area, ePhi, e2Phi, ePhi_bar, e2Phi_bar, atm_lgTb, atm_lgTeff, atm_lgflux, heat, T_atm,
        lnT, lnT0, epoch_boundaries, epoch_Mdots, enuc, dt, tsec, t_mon, Teff_mon, Qb_mon, Teff, Lsurf;
    DevBuf<double> tr_tsec, tr_dt, tr_Teff, tr_Lsurf, tr_Lnu, tr_Lnuc, tr_Mdot, tr_ext;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev_half = nullptr;
    // ordering between the two streams: face inputs / evolve inputs uploaded (copy stream), last compute issued (main)
    cudaEvent_t ev_face_in = nullptr, ev_evolve_in = nullptr, ev_busy = nullptr, ev_cell_in = nullptr, ev_Y_in = nullptr;
    char* h_stage = nullptr;   // pinned staging block for the small result arrays (download_results)
    size_t stage_bytes = 0;
    double ms[4] = {0, 0, 0, 0};   // cells, faces, evolve, EOS half of the cells (split pass)
    int launches = 0;
    // evolve launch plan: stars grouped by zones-per-lane class of the warp kernel (kernels_evolve_warp.cuh); the last
    // entry (Z = 0) is the CTA-per-star fallback for stars with more than 320 zones
    struct EvolveClass { int Z = 0, count = 0, max_nz = 0; DevBuf<int> list; };
    EvolveClass ev_class[DS_N_EVOLVE_CLASS + 1];
    // table sharing (dstar_batch_set_table_sharing), [0] = cell tables, [1] = face tables: owner star of each star's
    // tables, first table slot of each star in the compact arrays, zones of the owner stars (the work list of the pass)
    struct Share {
        bool on = false;
        std::vector<int> owner, tab0;
        size_t slots = 0;
        int n_work = 0;
        DevBuf<int> d_owner, d_tab0, d_work;
    } share[2];
    int evolve_mapping = 0;   // dstar_batch_set_evolve_mapping
    int evolve_minb = 0;      // > 0: CTAs per SM the evolve kernel is compiled for (experiments: DSTAR_EVOLVE_MINB): 0 = automatic, 1 = CTA per star, 2 = warp per star
};

namespace {
void fill_host_consts(DsHostConsts& h);
// per-species functions of Z for the ion-electron fits (eos.cuh: DsIonZ), evaluated here with the same dsmath
cudaError_t upload_ionz(DevBuf<DsIonZ>& d, std::vector<DsIonZ>& h, int ncharged, const double* Zion, cudaStream_t s) {
    DsHostConsts hc;
    fill_host_consts(hc);
    h.resize(ncharged);
    for (int i = 0; i < ncharged; ++i) h[i] = ds_ionz_compute(Zion[i], hc.ctf_fac);
    return d.upload(h.data(), ncharged, s);
}
void fill_host_consts(DsHostConsts& h) {
    using namespace dsc;
    h.theta_fac = 2.0 * dsm::pow((double)(4.0f / 9.0f) / pi, (double)(2.0f / 3.0f));
    h.ctf_fac = (double)(18.0f / 175.0f) * dsm::pow(12.0 / pi, 2.0 * onethird);
    h.xrs = finestructure * dsm::pow(9.0 * pi / 4.0, onethird);
    h.atf = (double)(54.0f / 175.0f) * dsm::pow(12.0 / pi, onethird) * finestructure;
    h.e1 = dsm::exp(1.0);
    h.half_log_6_over_pi = 0.5 * dsm::log(6.0 / pi);
    h.aei_fac = dsm::pow((double)(4.0f / 9.0f) / pi, onethird);
    h.befac = dsm::pow(finestructure / pi, 1.5);
    h.yfac = std::sqrt(4.0 * finestructure / pi);
    h.Tp_pre = hbar * std::sqrt(4.0 * pi * finestructure * hbar * clight / Melectron) / boltzmann;
}
// Coefficient-pass launch plan (kernels_micro.cuh, micro_*.cu).  Zone groups (of 128 lanes) per persistent CTA:
//   faces and the fused cell kernel 7 (896 threads, 72 registers; B200, 144 stars, fused cell pass: G=3 8.8 ms, 4 6.8,
//   5 5.7, 6 5.4, 7 5.2, 8 5.5 -- profiles/r01_notes.md), the EOS half 7, the neutrino half 8 (64 registers, no spill
//   frame).  DSTAR_MICRO_SPLIT=0 selects the fused cell kernel, DSTAR_MICRO_MINB / _G_EOS / _G_NU the group counts
//   (experiments; a count the library was not compiled for falls back to the default).
struct MicroPlan { int split, g_cell, g_eos, g_nu, g_face, blocks; };
const MicroPlan& micro_plan() {
    static const MicroPlan p = [] {
        auto geti = [](const char* k, int d) { const char* e = std::getenv(k); return e ? std::atoi(e) : d; };
        MicroPlan q;
        q.split = geti("DSTAR_MICRO_SPLIT", 1);
        q.g_cell = q.g_face = geti("DSTAR_MICRO_MINB", 7);
        q.g_face = geti("DSTAR_MICRO_G_FACE", q.g_face);
        q.g_eos = geti("DSTAR_MICRO_G_EOS", 7);
        q.g_nu = geti("DSTAR_MICRO_G_NU", 8);
        q.blocks = geti("DSTAR_MICRO_BLOCKS", 7);   // parts scheduled by 32-knot blocks instead of by zones: 1 EOS half, 2 neutrino half, 4 faces
        return q;
    }();
    return p;
}
DsEosCtrl eos_ctrl_from(const dstar_micro_ctrl* c) {
    DsEosCtrl e{175.0, 140.0, DS_R4(1.0e-9)};  // dStar_eos_def.f:58-62
    if (c) {
        if (c->eos_gamma_melt_pt > 0.0) e.Gamma_melt = c->eos_gamma_melt_pt;
        if (c->eos_rsi_melt_pt > 0.0) e.rsi_melt = c->eos_rsi_melt_pt;
        if (c->eos_nuclide_abundance_threshold > 0.0) e.Ythresh = c->eos_nuclide_abundance_threshold;
    }
    return e;
}
DsCondCtrl cond_ctrl_from(const dstar_micro_ctrl* c) {
    DsCondCtrl k{0, 0, 1, 1, 10.0, 9.0};  // conductivity_def.f:49-50
    if (c) {
        k.include_neutrons = c->use_neutron_conductivity;
        k.include_sf_phonons = c->use_superfluid_phonon_conductivity;
        k.ee_fmla = c->use_pcy_for_ee_scattering ? 2 : 1;
        k.eQ_fmla = c->use_page_for_eQ_scattering ? 2 : 1;
        if (c->rad_full_off_lgrho > 0.0) k.rad_full_off_lgrho = c->rad_full_off_lgrho;
        if (c->rad_full_on_lgrho > 0.0) k.rad_full_on_lgrho = c->rad_full_on_lgrho;
    }
    return k;
}
// CTA-per-star evolve kernel: block size from the largest star, CTAs per SM from the batch size (kernels_evolve.cuh)
cudaError_t launch_evolve_cta(const DsEvolveArgs& a, const int* list, int count, int max_nz, int sms, int minb_override, cudaStream_t s) {
    const int threads = ((max_nz + 31) / 32) * 32;
    // resident stars per SM (B200, 253-zone stars, ms for 148 / 256 / 512 / 1024 / 4096 stars -- profiles/r02_notes.md):
    //   1: 0.91 / 1.28 / 2.47 / 4.60 / 18.3    2: 1.15 / 1.32 / 2.03 / 3.80 / 16.2    4: 1.69 / 2.03 / 2.50 / 4.57 / 18.0
    // one star per SM (255 registers) wins up to two waves' worth of stars, then two per SM (128 registers); three and four
    // per SM spill too much to pay at any size.  Stars of 257-384 zones: one per SM throughout (1024 stars of 271 zones:
    // 6.84 against 7.83 ms).  (Long, uneven histories prefer two per SM earlier: accretion, 256 stars, 4.36 against 3.70 ms;
    // dstar_batch_set_evolve_occupancy / DSTAR_EVOLVE_MINB override the rule.)
    int minb = minb_override > 0 ? minb_override : (threads <= 256 ? (count <= 2 * sms ? 1 : 2) : 1);
    if (threads <= 256) {
        switch (minb) {
            case 1: ds_evolve_kernel<256, 1><<<count, threads, 0, s>>>(a, list); break;
            case 2: ds_evolve_kernel<256, 2><<<count, threads, 0, s>>>(a, list); break;
            case 3: ds_evolve_kernel<256, 3><<<count, threads, 0, s>>>(a, list); break;
            default: ds_evolve_kernel<256, 4><<<count, threads, 0, s>>>(a, list); break;
        }
    } else if (threads <= 384) {
        if (minb <= 1) ds_evolve_kernel<384, 1><<<count, threads, 0, s>>>(a, list);
        else ds_evolve_kernel<384, 2><<<count, threads, 0, s>>>(a, list);
    } else {
        ds_evolve_kernel<512, 1><<<count, threads, 0, s>>>(a, list);
    }
    return cudaGetLastError();
}
// warp-per-star evolve kernel for one zones-per-lane class: one star per 32-thread CTA, as many CTAs per SM as their
// shared memory allows (6 at Z = 8), so a small batch still spreads over all SMs
template <int Z>
cudaError_t launch_evolve_warp(const DsEvolveArgs& a, const int* list, int count, cudaStream_t s) {
    constexpr int S = 1;
    const size_t bytes = (DS_NTAB + (size_t)S * dsw::Lay<Z>::TOTAL) * sizeof(double);
    static bool once[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!once[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(ds_evolve_warp_kernel<Z, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        once[dev & 63] = true;
    }
    ds_evolve_warp_kernel<Z, S><<<(count + S - 1) / S, 32 * S, bytes, s>>>(a, list, count);
    return cudaGetLastError();
}
}  // namespace

static int g_force_redo = 0;

extern "C" {

const char* dstar_last_error(void) { return g_err.c_str(); }
void dstar_debug_force_redo(int on) { g_force_redo = on ? 1 : 0; }

// Local-memory window of the device.  Round 1's fused cell kernel kept a 0.7-0.8 KB spill frame per thread and its time
// depended on how the driver laid local memory out (5.4 vs 6.4 ms after NCCL had set cudaLimitStackSize to 1872 B), so
// the library raised the limit to 8 KB (2.4 GB of HBM).  With the round-2 kernels the dependence is gone (2 GPUs under
// NCCL: 7.892 ms per step with the device left alone, 7.894 with the 8 KB window; profiles/r02_notes.md), so the library
// no longer touches the limit unless asked to: DSTAR_STACK_LIMIT=<bytes> raises it (never lowers it).
static cudaError_t ensure_stack_limit() {
    static const size_t want = [] {
        const char* e = std::getenv("DSTAR_STACK_LIMIT");
        return e ? (size_t)std::strtoull(e, nullptr, 10) : (size_t)0;
    }();
    if (want == 0) return cudaSuccess;
    size_t cur = 0;
    cudaError_t e = cudaDeviceGetLimit(&cur, cudaLimitStackSize);
    if (e != cudaSuccess || cur >= want) return e;
    return cudaDeviceSetLimit(cudaLimitStackSize, want);   // synchronises the device
}

int dstar_ctx_create(int device, dstar_ctx** out) {
    if (!out) return fail(DSTAR_ERR_ARG, "ctx == NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(DSTAR_ERR_CUDA, "no CUDA device: dstar_b200 has no CPU fallback", e);
    if (device < 0 || device >= ndev) return fail(DSTAR_ERR_ARG, "device index out of range");
    CK(cudaSetDevice(device));
    CK(ensure_stack_limit());
    dstar_ctx* c = new dstar_ctx;
    c->device = device;
    CK(cudaGetDeviceProperties(&c->prop, device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    DsHostConsts& hc = c->hc;
    fill_host_consts(hc);
    CK(cudaMemcpyToSymbol(ds_hc, &hc, sizeof hc));
    // table knots (create_model.f:340-341), exp and log10 taken on the host
    double T[DS_NTAB], lgT[DS_NTAB];
    for (int i = 0; i < DS_NTAB; ++i) {
        c->h_tab_lnT[i] = 7.0 * dsc::ln10 + (10.0 - 7.0) * dsc::ln10 * (double)i / (double)(DS_NTAB - 1);
        T[i] = dsm::exp(c->h_tab_lnT[i]);
        lgT[i] = dsm::log10(T[i]);
    }
    CK(c->tab_lnT.upload(c->h_tab_lnT, DS_NTAB, c->stream));
    CK(c->tab_T.upload(T, DS_NTAB, c->stream));
    CK(c->tab_lgT.upload(lgT, DS_NTAB, c->stream));
    for (int i = 0; i < 3; ++i) { c->sf.loaded[i] = 0; c->sf.scale[i] = 1.0; c->sf.nv[i] = 0; }
    CK(cudaStreamSynchronize(c->stream));
    *out = c;
    return DSTAR_OK;
}

void dstar_ctx_destroy(dstar_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int dstar_device_info(dstar_ctx* c, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char name[256]) {
    if (!c) return fail(DSTAR_ERR_ARG, "ctx");
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (name) { std::strncpy(name, c->prop.name, 255); name[255] = 0; }
    return DSTAR_OK;
}

int dstar_ctx_set_helm(dstar_ctx* c, int imax, int jmax, const double* const a[21]) {
    if (!c || !a || imax < 2 || jmax < 2) return fail(DSTAR_ERR_ARG, "set_helm");
    CK(cudaSetDevice(c->device));
    // helm_alloc.f:150-165 geometry
    DsHelmDev& h = c->helm;
    h.imax = imax; h.jmax = jmax;
    const double logtlo = 3.0, logthi = 13.0, logdlo = -12.0, logdhi = 14.0;
    const double logtstp = (logthi - logtlo) / (double)(float)(jmax - 1);
    const double logdstp = (logdhi - logdlo) / (double)(float)(imax - 1);
    h.logtlo = logtlo; h.logtstpi = 1.0 / logtstp; h.logdlo = logdlo; h.logdstpi = 1.0 / logdstp;
    h.templo = 1.0e3; h.temphi = 1.0e13; h.logthi = logthi;
    h.denlo = 1.0e-12; h.denhi = 1.0e14;
    std::vector<double> t(jmax), d(imax);
    for (int j = 0; j < jmax; ++j) t[j] = dsm::exp((logtlo + j * logtstp) * dsc::ln10);
    for (int i = 0; i < imax; ++i) d[i] = dsm::exp((logdlo + i * logdstp) * dsc::ln10);
    // repack the 17 hot arrays into per-node records (see helm.cuh)
    static const int src[DS_HELM_REC] = {0, 2, 4, 1, 3, 5, 6, 7, 8, 9, 11, 10, 12, 13, 15, 14, 16};
    const size_t n = (size_t)imax * jmax;
    std::vector<double> node(n * DS_HELM_REC);
    for (size_t p = 0; p < n; ++p)
        for (int k = 0; k < DS_HELM_REC; ++k) node[p * DS_HELM_REC + k] = a[src[k]][p];
    CK(c->helm_node.upload(node.data(), node.size(), c->stream));
    CK(c->helm_t.upload(t.data(), t.size(), c->stream));
    CK(c->helm_d.upload(d.data(), d.size(), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    h.node = c->helm_node.p; h.t = c->helm_t.p; h.d = c->helm_d.p;
    c->helm_loaded = true;
    return DSTAR_OK;
}

int dstar_ctx_set_gaps(dstar_ctx* c, int type, int nv, double kF_min, double kF_max, const double* kF,
                       const double* f4) {
    if (!c || type < 0 || type > 2 || nv < 2 || !kF || !f4) return fail(DSTAR_ERR_ARG, "set_gaps");
    CK(cudaSetDevice(c->device));
    CK(c->sf_kF[type].upload(kF, nv, c->stream));
    CK(c->sf_f[type].upload(f4, (size_t)4 * nv, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->sf.loaded[type] = 1; c->sf.nv[type] = nv; c->sf.kF_min[type] = kF_min; c->sf.kF_max[type] = kF_max;
    c->sf.kF[type] = c->sf_kF[type].p; c->sf.f[type] = c->sf_f[type].p;
    return DSTAR_OK;
}

int dstar_ctx_set_sf_scale(dstar_ctx* c, const double s[3]) {
    if (!c || !s) return fail(DSTAR_ERR_ARG, "set_sf_scale");
    for (int i = 0; i < 3; ++i) c->sf.scale[i] = s[i];
    return DSTAR_OK;
}

// ------------------------------------------------------------ point evaluators
static int eval_crust_eos_impl(dstar_ctx* c, int n, const double* rho, const double* T, const double* ionic, int ncharged,
                         const double* Zion, const double* Aion, const double* Yion, const double* Tcs,
                         const double* chi_in, const double* eos_ctrl, double* res, int32_t* phase,
                         double* chi_out, int32_t* ierr, double* components) {
    if (!c || n <= 0) return fail(DSTAR_ERR_ARG, "eval_crust_eos");
    if (!c->helm_loaded) return fail(DSTAR_ERR_STATE, "HELM table not loaded (dstar_ctx_set_helm)");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf<double> d_rho, d_T, d_ion, d_Z, d_A, d_Y, d_Tc, d_chi, d_res, d_chio;
    DevBuf<int> d_ph, d_ie;
    CK(d_rho.upload(rho, n, s)); CK(d_T.upload(T, n, s)); CK(d_ion.upload(ionic, (size_t)11 * n, s));
    CK(d_Z.upload(Zion, ncharged, s)); CK(d_A.upload(Aion, ncharged, s));
    DevBuf<DsIonZ> d_ionz;
    std::vector<DsIonZ> h_ionz;
    CK(upload_ionz(d_ionz, h_ionz, ncharged, Zion, s));
    CK(d_Y.upload(Yion, (size_t)ncharged * n, s)); CK(d_Tc.upload(Tcs, (size_t)3 * n, s));
    if (chi_in) CK(d_chi.upload(chi_in, n, s));
    CK(d_res.alloc((size_t)15 * n)); CK(d_chio.alloc(n)); CK(d_ph.alloc(n)); CK(d_ie.alloc(n));
    DevBuf<double> d_comp;
    if (components) CK(d_comp.alloc((size_t)28 * n));
    DsEvalEosArgs a{};
    a.n = n; a.ncharged = ncharged; a.rho = d_rho.p; a.T = d_T.p; a.ionic = d_ion.p; a.Zion = d_Z.p; a.Aion = d_A.p;
    a.ionz = d_ionz.p;
    a.Yion = d_Y.p; a.Tcs = d_Tc.p; a.chi_in = chi_in ? d_chi.p : nullptr;
    a.eos = DsEosCtrl{175.0, 140.0, DS_R4(1.0e-9)};
    if (eos_ctrl) a.eos = DsEosCtrl{eos_ctrl[0], eos_ctrl[1], eos_ctrl[2]};
    a.res = d_res.p; a.phase = d_ph.p; a.chi_out = d_chio.p; a.ierr = d_ie.p; a.comps = components ? d_comp.p : nullptr;
    ds_eval_eos_kernel<<<(n + 63) / 64, 64, 0, s>>>(a, c->helm);
    CK(cudaGetLastError());
    CK(d_res.download(res, (size_t)15 * n, s));
    if (components) CK(d_comp.download(components, (size_t)28 * n, s));
    if (phase) CK(d_ph.download(phase, n, s));
    if (chi_out) CK(d_chio.download(chi_out, n, s));
    std::vector<int> ie(n);
    CK(d_ie.download(ie.data(), n, s));
    CK(cudaStreamSynchronize(s));
    int worst = 0;
    for (int i = 0; i < n; ++i) { if (ierr) ierr[i] = ie[i]; if (ie[i] != 0) worst = ie[i]; }
    return worst;
}

int dstar_eval_crust_eos(dstar_ctx* c, int n, const double* rho, const double* T, const double* ionic, int ncharged,
                         const double* Zion, const double* Aion, const double* Yion, const double* Tcs,
                         const double* chi_in, const double* eos_ctrl, double* res, int32_t* phase,
                         double* chi_out, int32_t* ierr) {
    return eval_crust_eos_impl(c, n, rho, T, ionic, ncharged, Zion, Aion, Yion, Tcs, chi_in, eos_ctrl, res, phase, chi_out,
                               ierr, nullptr);
}
int dstar_eval_crust_eos_components(dstar_ctx* c, int n, const double* rho, const double* T, const double* ionic,
                                    int ncharged, const double* Zion, const double* Aion, const double* Yion,
                                    const double* Tcs, const double* chi_in, const double* eos_ctrl, double* res,
                                    int32_t* phase, double* chi_out, int32_t* ierr, double* components) {
    return eval_crust_eos_impl(c, n, rho, T, ionic, ncharged, Zion, Aion, Yion, Tcs, chi_in, eos_ctrl, res, phase, chi_out,
                               ierr, components);
}

int dstar_crust_neutrino(dstar_ctx* c, int n, const double* rho, const double* T, const double* ionic,
                         const double* chi, const double* Tcn, const int32_t ch[5], double* eps) {
    if (!c || n <= 0) return fail(DSTAR_ERR_ARG, "crust_neutrino");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf<double> d_rho, d_T, d_ion, d_chi, d_Tc, d_eps;
    CK(d_rho.upload(rho, n, s)); CK(d_T.upload(T, n, s)); CK(d_ion.upload(ionic, (size_t)11 * n, s));
    CK(d_chi.upload(chi, n, s)); CK(d_Tc.upload(Tcn, n, s)); CK(d_eps.alloc((size_t)6 * n));
    DsEvalNuArgs a{};
    a.n = n; a.rho = d_rho.p; a.T = d_T.p; a.ionic = d_ion.p; a.chi = d_chi.p; a.Tcn = d_Tc.p; a.eps = d_eps.p;
    for (int i = 0; i < 5; ++i) a.channels[i] = ch ? ch[i] : 1;
    ds_eval_nu_kernel<<<(n + 63) / 64, 64, 0, s>>>(a);
    CK(cudaGetLastError());
    CK(d_eps.download(eps, (size_t)6 * n, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_conductivity(dstar_ctx* c, int n, const double* rho, const double* T, const double* chi,
                       const double* Gamma, const double* eta, const double* mu_e, const double* ionic,
                       const double* Tns, const double cc[6], double* K) {
    if (!c || n <= 0 || !cc) return fail(DSTAR_ERR_ARG, "conductivity");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf<double> d_rho, d_T, d_chi, d_G, d_eta, d_mu, d_ion, d_Tn, d_K;
    CK(d_rho.upload(rho, n, s)); CK(d_T.upload(T, n, s)); CK(d_chi.upload(chi, n, s)); CK(d_G.upload(Gamma, n, s));
    CK(d_eta.upload(eta, n, s)); CK(d_mu.upload(mu_e, n, s)); CK(d_ion.upload(ionic, (size_t)11 * n, s));
    CK(d_Tn.upload(Tns, n, s)); CK(d_K.alloc((size_t)10 * n));
    DsEvalCondArgs a{};
    a.n = n; a.rho = d_rho.p; a.T = d_T.p; a.chi = d_chi.p; a.Gamma = d_G.p; a.eta = d_eta.p; a.mu_e = d_mu.p;
    a.ionic = d_ion.p; a.Tns = d_Tn.p; a.K = d_K.p;
    a.cond = DsCondCtrl{cc[0] != 0, cc[1] != 0, (int)cc[2], (int)cc[3], cc[4], cc[5]};
    ds_eval_cond_kernel<<<(n + 31) / 32, 32, 0, s>>>(a);
    CK(cudaGetLastError());
    CK(d_K.download(K, (size_t)10 * n, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_sf_get_results(dstar_ctx* c, int n, const double* kFp, const double* kFn, double* Tc) {
    if (!c || n <= 0) return fail(DSTAR_ERR_ARG, "sf_get_results");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf<double> a, b, o;
    CK(a.upload(kFp, n, s)); CK(b.upload(kFn, n, s)); CK(o.alloc((size_t)3 * n));
    ds_sf_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, a.p, b.p, o.p, c->sf);
    CK(cudaGetLastError());
    CK(o.download(Tc, (size_t)3 * n, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_crust_composition_info(dstar_ctx* c, int nv, int nisos, const double* lgP_tab, const double* Ycoef,
                                 const double* Ziso, const double* Aiso, int nz, const double* lgP, double* ionic,
                                 double* Yion, int32_t* ncharged) {
    if (!c || nv < 2 || nisos < 1 || nz < 1) return fail(DSTAR_ERR_ARG, "crust_composition_info");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    int nch = 0;
    std::vector<double> z53(nisos), z73(nisos), z52(nisos), zz1(nisos);
    for (int i = 0; i < nisos; ++i) {
        if (Ziso[i] > 0.0 && Aiso[i] > 0.0) ++nch;
        z53[i] = dsm::pow(Ziso[i], 5.0 / 3.0); z73[i] = dsm::pow(Ziso[i], 7.0 / 3.0);
        z52[i] = dsm::pow(Ziso[i], 2.5); zz1[i] = Ziso[i] * dsm::pow(Ziso[i] + 1.0, 1.5);
    }
    DevBuf<double> d_tab, d_Y, d_Z, d_A, d_53, d_73, d_52, d_zz, d_lgP, d_ion, d_Yion;
    CK(d_tab.upload(lgP_tab, nv, s)); CK(d_Y.upload(Ycoef, (size_t)4 * nv * nisos, s));
    CK(d_Z.upload(Ziso, nisos, s)); CK(d_A.upload(Aiso, nisos, s)); CK(d_53.upload(z53.data(), nisos, s));
    CK(d_73.upload(z73.data(), nisos, s)); CK(d_52.upload(z52.data(), nisos, s)); CK(d_zz.upload(zz1.data(), nisos, s));
    CK(d_lgP.upload(lgP, nz, s)); CK(d_ion.alloc((size_t)11 * nz)); CK(d_Yion.alloc((size_t)nch * nz));
    DsCompArgs a{nz, nv, nisos, nch, d_tab.p, d_Y.p, d_Z.p, d_A.p, d_53.p, d_73.p, d_52.p, d_zz.p, d_lgP.p, d_ion.p,
                 d_Yion.p};
    ds_crust_composition_kernel<<<(nz + 127) / 128, 128, 0, s>>>(a);
    CK(cudaGetLastError());
    CK(d_ion.download(ionic, (size_t)11 * nz, s));
    CK(d_Yion.download(Yion, (size_t)nch * nz, s));
    CK(cudaStreamSynchronize(s));
    if (ncharged) *ncharged = nch;
    return DSTAR_OK;
}

// do_get_bc09_Teff (dStar_atm/private/bc09.f:45-95) for n_tab (grav, Plight, Pb) triples at once: one thread per
// (table, knot), kernels_atm.cuh.  lgflux = 4 lgTeff + log10(sigma_SB) (bc09.f:88) is filled here.
int dstar_atm_bc09_Teff(dstar_ctx* c, int n_tab, const double* grav, const double* Plight, const double* Pb, int nv,
                        const double* lgTeff, double* lgTb, double* lgflux, int32_t* ierr, double* info) {
    if (!c || n_tab < 1 || nv < 1 || !grav || !Plight || !Pb || !lgTeff || !lgTb) return fail(DSTAR_ERR_ARG, "atm_bc09_Teff");
    if (!c->helm_loaded) return fail(DSTAR_ERR_STATE, "HELM table not loaded (dstar_ctx_set_helm)");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t tot = (size_t)n_tab * nv;
    DevBuf<double> d_g, d_pl, d_pb, d_te, d_tb, d_info;
    DevBuf<int> d_ie;
    CK(d_g.upload(grav, n_tab, s)); CK(d_pl.upload(Plight, n_tab, s)); CK(d_pb.upload(Pb, n_tab, s));
    CK(d_te.upload(lgTeff, nv, s)); CK(d_tb.alloc(tot)); CK(d_ie.alloc(tot));
    if (info) CK(d_info.alloc(4 * tot));
    DsHostConsts hc;
    fill_host_consts(hc);
    CK(ds_bc09_launch(c->helm, hc, n_tab, nv, d_g.p, d_pl.p, d_pb.p, d_te.p, d_tb.p, info ? d_info.p : nullptr, d_ie.p, s));
    std::vector<int> h_ie(tot);
    CK(d_tb.download(lgTb, tot, s));
    CK(d_ie.download(h_ie.data(), tot, s));
    if (info) CK(d_info.download(info, 4 * tot, s));
    CK(cudaStreamSynchronize(s));
    int worst = 0;
    for (int t = 0; t < n_tab; ++t) {
        int e = 0;
        for (int i = nv - 1; i >= 0; --i)   // the reference stops at the first failing knot, hottest first (bc09.f:77-87)
            if (h_ie[(size_t)t * nv + i] != 0) { e = h_ie[(size_t)t * nv + i]; break; }
        if (ierr) ierr[t] = e;
        if (e != 0) worst = e;
    }
    if (lgflux) {
        const double lg_sigma = dsm::log10(dsc::sigma_SB);
        for (int t = 0; t < n_tab; ++t)
            for (int i = 0; i < nv; ++i) lgflux[(size_t)t * nv + i] = 4.0 * lgTeff[i] + lg_sigma;
    }
    if (worst != 0) return fail(worst, "bc09: envelope integration failed for at least one table (see ierr)");
    return DSTAR_OK;
}

// ------------------------------------------------------------------ batches
int dstar_batch_create(dstar_ctx* c, int n_star, const int32_t* zone_offset, int ncharged, dstar_batch** out) {
    if (!c || !out || n_star < 1 || !zone_offset || ncharged < 1) return fail(DSTAR_ERR_ARG, "batch_create");
    CK(cudaSetDevice(c->device));
    dstar_batch* b = new dstar_batch;
    b->ctx = c; b->n_star = n_star; b->ncharged = ncharged;
    b->h_zone_offset.assign(zone_offset, zone_offset + n_star + 1);
    b->total = zone_offset[n_star];
    std::vector<int> zs(b->total);
    for (int s = 0; s < n_star; ++s) {
        int nz = zone_offset[s + 1] - zone_offset[s];
        if (nz < 3 || nz > DS_MAXZ) { delete b; return fail(DSTAR_ERR_ARG, "each star needs 3 <= nz <= 512 zones"); }
        b->max_nz = std::max(b->max_nz, nz);
        for (int z = zone_offset[s]; z < zone_offset[s + 1]; ++z) zs[z] = s;
    }
    cudaStream_t st = c->stream;
    CK(b->zone_offset.upload(b->h_zone_offset.data(), n_star + 1, st));
    CK(b->zone_star.upload(zs.data(), b->total, st));
    CK(b->status.alloc(n_star)); CK(cudaMemsetAsync(b->status.p, 0, n_star * sizeof(int), st));
    CK(b->redo.alloc(2 * (size_t)b->total)); CK(b->redo_count.alloc(2));   // [0] EOS half / fused pass, [1] neutrino half
    CK(b->rho_bar1.alloc(n_star));
    {   // evolve launch plan (see dstar_batch::ev_class)
        std::vector<int> lists[DS_N_EVOLVE_CLASS + 1];
        for (int s = 0; s < n_star; ++s) {
            const int nz = zone_offset[s + 1] - zone_offset[s];
            int cls = DS_N_EVOLVE_CLASS;
            for (int q = 0; q < DS_N_EVOLVE_CLASS; ++q)
                if (nz <= 32 * DS_EVOLVE_Z[q]) { cls = q; break; }
            lists[cls].push_back(s);
            b->ev_class[cls].max_nz = std::max(b->ev_class[cls].max_nz, nz);
        }
        for (int q = 0; q <= DS_N_EVOLVE_CLASS; ++q) {
            b->ev_class[q].Z = q < DS_N_EVOLVE_CLASS ? DS_EVOLVE_Z[q] : 0;
            b->ev_class[q].count = (int)lists[q].size();
            if (!lists[q].empty()) CK(b->ev_class[q].list.upload(lists[q].data(), lists[q].size(), st));
        }
    }
    { static const int v = [] { const char* e = std::getenv("DSTAR_EVOLVE_MINB"); return e ? std::atoi(e) : 0; }(); b->evolve_minb = v; }
    CK(cudaEventCreate(&b->ev0)); CK(cudaEventCreate(&b->ev1)); CK(cudaEventCreate(&b->ev_half));
    CK(cudaEventCreate(&b->ev2)); CK(cudaEventCreate(&b->ev3));
    CK(cudaEventCreateWithFlags(&b->ev_face_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_evolve_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_busy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_cell_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_Y_in, cudaEventDisableTiming));
    CK(cudaStreamSynchronize(st));
    *out = b;
    return DSTAR_OK;
}

void dstar_batch_destroy(dstar_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->copy_stream); cudaStreamSynchronize(b->ctx->stream);
    if (b->ev_face_in) cudaEventDestroy(b->ev_face_in);
    if (b->ev_evolve_in) cudaEventDestroy(b->ev_evolve_in);
    if (b->ev_busy) cudaEventDestroy(b->ev_busy);
    if (b->ev_cell_in) cudaEventDestroy(b->ev_cell_in);
    if (b->ev_Y_in) cudaEventDestroy(b->ev_Y_in);
    if (b->h_stage) cudaFreeHost(b->h_stage);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev_half) cudaEventDestroy(b->ev_half);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->ev2) cudaEventDestroy(b->ev2);
    if (b->ev3) cudaEventDestroy(b->ev3);
    delete b;
}

int dstar_batch_upload_structure(dstar_batch* b, const double* rho, const double* rho_bar, const double* P,
                                 const double* P_bar, const double* dm, const double* ionic, const double* ionic_bar,
                                 const double* Yion, const double* Yion_bar, const double* Zion, const double* Aion,
                                 const double* Tc_cell, const double* Tc_face) {
    if (!b || !rho || !rho_bar || !P || !P_bar || !dm || !ionic || !ionic_bar || !Yion || !Yion_bar || !Zion || !Aion)
        return fail(DSTAR_ERR_ARG, "batch_upload_structure");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t n = b->total;
    // Inputs of the cell pass go on the compute stream; what only the face pass needs goes on the copy stream and
    // overlaps the cell pass.  The copy stream first waits for the compute issued so far (no overwrite of buffers a
    // running kernel still reads).  With pinned host arrays these copies are asynchronous: the arrays must stay
    // untouched until the next call that synchronises (micro_tables, evolve, download_*).
    cudaStream_t cs = b->ctx->copy_stream;
    CK(cudaStreamWaitEvent(cs, b->ev_busy, 0));
    CK(b->rho.upload(rho, n, s)); CK(b->P.upload(P, n, s));
    CK(b->P_bar.upload(P_bar, n, s)); CK(b->dm.upload(dm, n, s)); CK(b->ionic.upload(ionic, 11 * n, s));
    CK(b->Zion.upload(Zion, b->ncharged, s));
    CK(b->Aion.upload(Aion, b->ncharged, s));
    CK(upload_ionz(b->ionz, b->h_ionz, b->ncharged, Zion, s));
    b->have_Tc_cell = Tc_cell != nullptr; b->have_Tc_face = Tc_face != nullptr;
    if (Tc_cell) CK(b->Tc_cell.upload(Tc_cell, 3 * n, s));
    // rho_bar(1) starts from the uploaded value of each star's first face
    b->h_rb1.resize(b->n_star);
    for (int i = 0; i < b->n_star; ++i) b->h_rb1[i] = rho_bar[b->h_zone_offset[i]];
    CK(b->rho_bar1.upload(b->h_rb1.data(), b->n_star, s));
    // the copy stream starts only when the first kernels' inputs have landed: copies of both streams share the host
    // link, and interleaved they delayed the start of the cell pass by ~0.4 ms (256 stars); now they run under the
    // kernels.  The abundances (half of the cell inputs) go first on the copy stream: the neutrino kernel, which runs
    // first, does not read them; the EOS kernel waits for ev_Y_in.
    CK(cudaEventRecord(b->ev_cell_in, s));
    CK(cudaStreamWaitEvent(cs, b->ev_cell_in, 0));
    CK(b->Yion.upload(Yion, b->ncharged * n, cs));
    CK(cudaEventRecord(b->ev_Y_in, cs));
    CK(b->rho_bar.upload(rho_bar, n, cs)); CK(b->ionic_bar.upload(ionic_bar, 11 * n, cs));
    CK(b->Yion_bar.upload(Yion_bar, b->ncharged * n, cs));
    if (Tc_face) CK(b->Tc_face.upload(Tc_face, 3 * n, cs));
    CK(cudaEventRecord(b->ev_face_in, cs));
    b->have_structure = true;
    return DSTAR_OK;
}

int dstar_batch_set_Tc_face(dstar_batch* b, const double* Tc_face) {
    if (!b || !Tc_face) return fail(DSTAR_ERR_ARG, "set_Tc_face");
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamWaitEvent(b->ctx->stream, b->ev_face_in, 0));
    CK(b->Tc_face.upload(Tc_face, (size_t)3 * b->total, b->ctx->stream));
    CK(cudaStreamSynchronize(b->ctx->stream));
    b->have_Tc_face = true;
    return DSTAR_OK;
}
int dstar_batch_set_ionic_bar(dstar_batch* b, const double* ionic_bar) {
    if (!b || !ionic_bar) return fail(DSTAR_ERR_ARG, "set_ionic_bar");
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamWaitEvent(b->ctx->stream, b->ev_face_in, 0));
    CK(b->ionic_bar.upload(ionic_bar, (size_t)11 * b->total, b->ctx->stream));
    CK(cudaStreamSynchronize(b->ctx->stream));
    return DSTAR_OK;
}
int dstar_batch_get_rho_bar1(dstar_batch* b, double* rho_bar1) {
    if (!b || !rho_bar1) return fail(DSTAR_ERR_ARG, "get_rho_bar1");
    CK(cudaSetDevice(b->ctx->device));
    CK(b->rho_bar1.download(rho_bar1, b->n_star, b->ctx->stream));
    CK(cudaStreamSynchronize(b->ctx->stream));
    return DSTAR_OK;
}

int dstar_batch_micro_tables(dstar_batch* b, const dstar_micro_ctrl* ctrl, int which) {
    if (!b || which < 1 || which > 3) return fail(DSTAR_ERR_ARG, "batch_micro_tables");
    if (!b->have_structure) return fail(DSTAR_ERR_STATE, "structure not uploaded");
    dstar_ctx* c = b->ctx;
    if (!c->helm_loaded) return fail(DSTAR_ERR_STATE, "HELM table not loaded (dstar_ctx_set_helm)");
    CK(cudaSetDevice(c->device));
    CK(ensure_stack_limit());
    cudaStream_t s = c->stream;
    const size_t csz = (size_t)4 * DS_NTAB * (b->share[0].on ? b->share[0].slots : (size_t)b->total);
    const size_t fsz = (size_t)4 * DS_NTAB * (b->share[1].on ? b->share[1].slots : (size_t)b->total);
    if ((which & 1) && b->tab_lnCp.n != csz) { CK(b->tab_lnCp.alloc(csz)); CK(b->tab_lnGamma.alloc(csz)); CK(b->tab_lnEnu.alloc(csz)); }
    if ((which & 2) && b->tab_lnK.n != fsz) CK(b->tab_lnK.alloc(fsz));
    DsMicroArgs a{};
    a.total_zones = b->total; a.ncharged = b->ncharged; a.zone_star = b->zone_star.p; a.zone_offset = b->zone_offset.p;
    a.rho = b->rho.p; a.rho_bar = b->rho_bar.p; a.P = b->P.p; a.P_bar = b->P_bar.p; a.dm = b->dm.p;
    a.ionic = b->ionic.p; a.ionic_bar = b->ionic_bar.p; a.Yion = b->Yion.p; a.Yion_bar = b->Yion_bar.p;
    a.Zion = b->Zion.p; a.Aion = b->Aion.p; a.ionz = b->ionz.p;
    if (!b->aux.p) CK(b->aux.alloc((size_t)DS_AUX * b->total));
    a.aux = b->aux.p;
    a.Tc_cell = b->have_Tc_cell ? b->Tc_cell.p : nullptr;
    a.Tc_face = b->have_Tc_face ? b->Tc_face.p : nullptr;
    a.tab_lnT = c->tab_lnT.p; a.tab_T = c->tab_T.p; a.tab_lgT = c->tab_lgT.p;
    a.eos = eos_ctrl_from(ctrl); a.cond = cond_ctrl_from(ctrl);
    const int defch[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 5; ++i) a.nu_channels[i] = defch[i];
    if (ctrl) {
        a.nu_channels[0] = ctrl->use_crust_nu_pair; a.nu_channels[1] = ctrl->use_crust_nu_photo;
        a.nu_channels[2] = ctrl->use_crust_nu_plasma; a.nu_channels[3] = ctrl->use_crust_nu_bremsstrahlung;
        a.nu_channels[4] = ctrl->use_crust_nu_pbf;
        a.turn_on_shell_Urca = ctrl->turn_on_shell_Urca;
        a.shell_Urca_luminosity_coeff = ctrl->shell_Urca_luminosity_coeff; a.lgP_shell_Urca = ctrl->lgP_shell_Urca;
    }
    a.rho_bar1 = b->rho_bar1.p; a.status = b->status.p;
    a.tab_lnCp = b->tab_lnCp.p; a.tab_lnGamma = b->tab_lnGamma.p; a.tab_lnEnu = b->tab_lnEnu.p; a.tab_lnK = b->tab_lnK.p;
    b->launches = 0;
    if (which & 1) {
        CK(cudaEventRecord(b->ev0, s));
        {   // one persistent CTA per SM, G zone-groups of 128 lanes in lockstep (see kernels_micro.cuh)
            const int sms = c->prop.multiProcessorCount;
            a.redo = b->redo.p; a.redo_count = b->redo_count.p;
            const dstar_batch::Share& sh = b->share[0];
            a.work_list = sh.on ? sh.d_work.p : nullptr; a.n_work = sh.n_work; a.tab_zone0 = sh.on ? sh.d_tab0.p : nullptr;
            ds_zone_aux_kernel<false><<<(b->total + 127) / 128, 128, 0, s>>>(a, c->helm, c->sf);
            CK(cudaMemsetAsync(b->redo_count.p, 0, 2 * sizeof(int), s));
            a.force_redo = g_force_redo;   // test hook (dstar_debug_force_redo)
            const int n_items = sh.on ? sh.n_work : b->total;
            const MicroPlan& mp = micro_plan();
            // branch-free pass, then the plain pass over the zones it flagged (normally none); the two halves of the
            // split pass keep separate lists
            if (mp.blocks && !b->blk_list.p) { CK(b->blk_list.alloc((size_t)20 * b->total)); CK(b->blk_count.alloc(5 + 3 * DS_BLK_NKEY)); }
            if (mp.blocks) {
                CK(cudaMemsetAsync(b->blk_count.p, 0, (5 + 3 * DS_BLK_NKEY) * sizeof(int), s));
                a.blk_list = b->blk_list.p; a.blk_count = b->blk_count.p;
            }
            if (mp.split) {   // the neutrino half first: it needs no abundances, which may still be arriving (ev_Y_in)
                a.redo = b->redo.p + b->total; a.redo_count = b->redo_count.p + 1;
                if (mp.blocks & 2) {
                    a.blk_redo = b->blk_list.p + (size_t)16 * b->total; a.blk_redo_count = b->blk_count.p + 4 + 3 * DS_BLK_NKEY;
                    CK(ds_launch_micro_nu_blocks(sms, n_items, s, a, c->helm, c->hc));
                    ++b->launches;
                } else {
                    CK(ds_launch_micro_nu(mp.g_nu, true, sms, n_items, s, a, c->helm, c->sf, c->hc));
                    CK(ds_launch_micro_nu(mp.g_nu, false, sms, n_items, s, a, c->helm, c->sf, c->hc));
                }
                CK(cudaEventRecord(b->ev_half, s));
                CK(cudaStreamWaitEvent(s, b->ev_Y_in, 0));
                a.redo = b->redo.p; a.redo_count = b->redo_count.p;
                if (mp.blocks & 1) {
                    CK(ds_launch_micro_eos_blocks(sms, n_items, s, a, c->helm, c->hc));
                    b->launches += 6;
                } else {
                    CK(ds_launch_micro_eos(mp.g_eos, true, sms, n_items, s, a, c->helm, c->sf, c->hc));
                    CK(ds_launch_micro_eos(mp.g_eos, false, sms, n_items, s, a, c->helm, c->sf, c->hc));
                    b->launches += 2;
                }
            } else {
                CK(cudaStreamWaitEvent(s, b->ev_Y_in, 0));
                CK(ds_launch_micro_cell(mp.g_cell, true, sms, n_items, s, a, c->helm, c->sf, c->hc));
                CK(ds_launch_micro_cell(mp.g_cell, false, sms, n_items, s, a, c->helm, c->sf, c->hc));
            }
            if (sh.on) { ds_share_rho_bar1_kernel<<<(b->n_star + 127) / 128, 128, 0, s>>>(b->n_star, sh.d_owner.p, b->rho_bar1.p); ++b->launches; }
            b->launches += 3;   // zone pre-pass, branch-free pass, plain pass over flagged zones (+ 2 more when split)
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(b->ev1, s));   // read back after the face pass has been queued: no idle gap between them
    }
    if (which & 2) {
        CK(cudaEventRecord(b->ev2, s));
        {   // one persistent CTA per SM, G zone-groups of 128 lanes in lockstep (see kernels_micro.cuh)
            const int sms = c->prop.multiProcessorCount;
            a.redo = nullptr; a.redo_count = nullptr;
            const dstar_batch::Share& sh = b->share[1];
            a.work_list = sh.on ? sh.d_work.p : nullptr; a.n_work = sh.n_work; a.tab_zone0 = sh.on ? sh.d_tab0.p : nullptr;
            CK(cudaStreamWaitEvent(s, b->ev_face_in, 0));   // face inputs may still be arriving on the copy stream
            ds_zone_aux_kernel<true><<<(b->total + 127) / 128, 128, 0, s>>>(a, c->helm, c->sf);
            if (micro_plan().blocks & 4) {
                CK(ds_launch_micro_face_blocks(sms, sh.on ? sh.n_work : b->total, s, a, c->helm, c->hc));
                ++b->launches;
            } else {
                CK(ds_launch_micro_face(micro_plan().g_face, false, sms, sh.on ? sh.n_work : b->total, s, a, c->helm, c->sf, c->hc));
            }
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(b->ev3, s));
        CK(cudaEventSynchronize(b->ev3));
        float ms = 0; CK(cudaEventElapsedTime(&ms, b->ev2, b->ev3)); b->ms[1] = ms;
        b->launches += 2;   // zone pre-pass, face pass
    }
    if (which & 1) {
        CK(cudaEventSynchronize(b->ev1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1)); b->ms[0] = ms;
        b->ms[3] = 0.0;
        if (micro_plan().split) { CK(cudaEventElapsedTime(&ms, b->ev_half, b->ev1)); b->ms[3] = ms; }   // the EOS half runs second
    }
    CK(cudaEventRecord(b->ev_busy, s));
    b->have_tables = true;
    return DSTAR_OK;
}

int dstar_batch_set_table_sharing(dstar_batch* b, int which, const int32_t* owner) {
    if (!b || (which != 1 && which != 2)) return fail(DSTAR_ERR_ARG, "batch_set_table_sharing: which = 1 (cell tables) or 2 (face tables)");
    dstar_batch::Share& sh = b->share[which - 1];
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    b->have_tables = false;
    if (!owner) { sh.on = false; return DSTAR_OK; }
    const int n = b->n_star;
    sh.owner.assign(owner, owner + n);
    sh.tab0.assign(n, -1);
    std::vector<int> work;
    size_t slots = 0;
    for (int i = 0; i < n; ++i) {
        const int o = owner[i];
        if (o < 0 || o >= n || owner[o] != o) return fail(DSTAR_ERR_ARG, "table sharing: owner[i] must name a star that owns its tables");
        const int nz = b->h_zone_offset[i + 1] - b->h_zone_offset[i];
        if (b->h_zone_offset[o + 1] - b->h_zone_offset[o] != nz) return fail(DSTAR_ERR_ARG, "table sharing: stars with different zone counts");
        if (o == i) {
            sh.tab0[i] = (int)slots;
            slots += nz;
            for (int k = 0; k < nz; ++k) work.push_back(b->h_zone_offset[i] + k);
        }
    }
    for (int i = 0; i < n; ++i) {
        if (owner[i] > i) return fail(DSTAR_ERR_ARG, "table sharing: the owner must be the first star of its group");
        sh.tab0[i] = sh.tab0[owner[i]];
    }
    sh.slots = slots; sh.n_work = (int)work.size();
    CK(sh.d_owner.upload(sh.owner.data(), n, s)); CK(sh.d_tab0.upload(sh.tab0.data(), n, s)); CK(sh.d_work.upload(work.data(), work.size(), s));
    CK(cudaStreamSynchronize(s));
    sh.on = true;
    return DSTAR_OK;
}

int dstar_batch_download_tables(dstar_batch* b, double* tab_lnT, double* tab_lnCp, double* tab_lnGamma,
                                double* tab_lnEnu, double* tab_lnK, double* rho_bar1, int32_t* status) {
    if (!b) return fail(DSTAR_ERR_ARG, "batch_download_tables");
    if (!b->have_tables) return fail(DSTAR_ERR_STATE, "tables not computed");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t tsz = (size_t)4 * DS_NTAB * b->total;
    if (tab_lnT) std::memcpy(tab_lnT, b->ctx->h_tab_lnT, sizeof(double) * DS_NTAB);
    {   // with shared tables the device holds one copy per owner; the caller gets every star's (4*128, nz) block
        struct Item { double* host; const DevBuf<double>* dev; int which; };
        const Item items[4] = {{tab_lnCp, &b->tab_lnCp, 0}, {tab_lnGamma, &b->tab_lnGamma, 0}, {tab_lnEnu, &b->tab_lnEnu, 0}, {tab_lnK, &b->tab_lnK, 1}};
        std::vector<double> tmp;
        for (const Item& it : items) {
            if (!it.host) continue;
            const dstar_batch::Share& sh = b->share[it.which];
            if (!sh.on) { CK(it.dev->download(it.host, tsz, s)); continue; }
            tmp.resize((size_t)4 * DS_NTAB * sh.slots);
            CK(it.dev->download(tmp.data(), tmp.size(), s));
            CK(cudaStreamSynchronize(s));
            for (int i = 0; i < b->n_star; ++i) {
                const size_t nz = b->h_zone_offset[i + 1] - b->h_zone_offset[i];
                std::memcpy(it.host + (size_t)4 * DS_NTAB * b->h_zone_offset[i], tmp.data() + (size_t)4 * DS_NTAB * sh.tab0[i],
                            sizeof(double) * 4 * DS_NTAB * nz);
            }
        }
    }
    if (rho_bar1) CK(b->rho_bar1.download(rho_bar1, b->n_star, s));
    if (status) CK(b->status.download(status, b->n_star, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_batch_upload_tables(dstar_batch* b, const double* tab_lnT, const double* tab_lnCp, const double* tab_lnGamma,
                              const double* tab_lnEnu, const double* tab_lnK) {
    if (!b || !tab_lnCp || !tab_lnGamma || !tab_lnEnu || !tab_lnK) return fail(DSTAR_ERR_ARG, "batch_upload_tables");
    (void)tab_lnT;
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t tsz = (size_t)4 * DS_NTAB * b->total;
    b->share[0].on = false; b->share[1].on = false;   // the caller brings every star's own tables
    CK(b->tab_lnCp.upload(tab_lnCp, tsz, s)); CK(b->tab_lnGamma.upload(tab_lnGamma, tsz, s));
    CK(b->tab_lnEnu.upload(tab_lnEnu, tsz, s)); CK(b->tab_lnK.upload(tab_lnK, tsz, s));
    CK(cudaStreamSynchronize(s));
    b->have_tables = true;
    return DSTAR_OK;
}

int dstar_batch_upload_evolve(dstar_batch* b, const double* dm_bar, const double* area, const double* ePhi,
                              const double* e2Phi, const double* ePhi_bar, const double* e2Phi_bar, int n_atm,
                              int atm_nv, const double* atm_lgTb, const double* atm_lgTeff, const double* atm_lgflux,
                              const int32_t* atm_index, const double* heat, const double* T_atm, const double* lnT0,
                              int n_epoch, const double* epoch_boundaries, const double* epoch_Mdots,
                              const double* enuc_per_Mdot) {
    if (!b || !dm_bar || !area || !ePhi || !e2Phi || !ePhi_bar || !e2Phi_bar || n_atm < 1 || atm_nv < 2 || !atm_lgTb ||
        !atm_lgTeff || !atm_lgflux || !atm_index || !heat || !T_atm || !lnT0 || n_epoch < 1 || !epoch_boundaries ||
        !epoch_Mdots)
        return fail(DSTAR_ERR_ARG, "batch_upload_evolve");
    if (!b->have_structure) return fail(DSTAR_ERR_STATE, "structure not uploaded");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->copy_stream;   // only the evolve kernel reads these: overlap the coefficient pass
    CK(cudaStreamWaitEvent(s, b->ev_busy, 0));
    const size_t n = b->total, ns = b->n_star;
    CK(b->dm_bar.upload(dm_bar, n, s)); CK(b->area.upload(area, n, s)); CK(b->ePhi.upload(ePhi, n, s));
    CK(b->e2Phi.upload(e2Phi, n, s)); CK(b->ePhi_bar.upload(ePhi_bar, n, s)); CK(b->e2Phi_bar.upload(e2Phi_bar, n, s));
    CK(b->atm_lgTb.upload(atm_lgTb, (size_t)n_atm * atm_nv, s));
    CK(b->atm_lgTeff.upload(atm_lgTeff, (size_t)n_atm * 4 * atm_nv, s));
    CK(b->atm_lgflux.upload(atm_lgflux, (size_t)n_atm * 4 * atm_nv, s));
    CK(b->atm_index.upload(atm_index, ns, s)); CK(b->heat.upload(heat, 9 * ns, s)); CK(b->T_atm.upload(T_atm, ns, s));
    CK(b->lnT.upload(lnT0, n, s));
    CK(b->lnT0.upload(lnT0, n, s));
    CK(b->epoch_boundaries.upload(epoch_boundaries, (size_t)(n_epoch + 1) * ns, s));
    CK(b->epoch_Mdots.upload(epoch_Mdots, (size_t)n_epoch * ns, s));
    b->have_enuc = enuc_per_Mdot != nullptr;
    if (enuc_per_Mdot) CK(b->enuc.upload(enuc_per_Mdot, n, s));
    b->n_epoch = n_epoch; b->n_atm = n_atm; b->atm_nv = atm_nv;
    CK(b->dt.ensure(ns)); CK(b->tsec.ensure(ns)); CK(b->model.ensure(ns));
    CK(cudaMemsetAsync(b->dt.p, 0, ns * sizeof(double), s));   // NScool_init.f:59-62
    CK(cudaMemsetAsync(b->tsec.p, 0, ns * sizeof(double), s));
    CK(cudaMemsetAsync(b->model.p, 0, ns * sizeof(int), s));
    CK(b->t_mon.ensure((size_t)n_epoch * ns)); CK(b->Teff_mon.ensure((size_t)n_epoch * ns));
    CK(b->Qb_mon.ensure((size_t)n_epoch * ns)); CK(b->idid.ensure((size_t)n_epoch * ns));
    CK(b->counters.ensure(7 * ns)); CK(b->Teff.ensure(ns)); CK(b->Lsurf.ensure(ns));
    CK(cudaEventRecord(b->ev_evolve_in, s));
    b->have_evolve = true;
    return DSTAR_OK;
}

int dstar_batch_reset_state(dstar_batch* b) {
    if (!b) return fail(DSTAR_ERR_ARG, "batch_reset_state");
    if (!b->have_evolve) return fail(DSTAR_ERR_STATE, "evolve inputs missing");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t ns = b->n_star;
    CK(cudaStreamWaitEvent(s, b->ev_evolve_in, 0));
    CK(cudaMemcpyAsync(b->lnT.p, b->lnT0.p, (size_t)b->total * sizeof(double), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemsetAsync(b->dt.p, 0, ns * sizeof(double), s));
    CK(cudaMemsetAsync(b->tsec.p, 0, ns * sizeof(double), s));
    CK(cudaMemsetAsync(b->model.p, 0, ns * sizeof(int), s));
    return DSTAR_OK;
}

int dstar_batch_set_evolve_occupancy(dstar_batch* b, int ctas_per_sm) {
    if (!b || ctas_per_sm < 0 || ctas_per_sm > 4) return fail(DSTAR_ERR_ARG, "evolve occupancy: 0 (automatic) .. 4 CTAs per SM");
    b->evolve_minb = ctas_per_sm;
    return DSTAR_OK;
}

static DsEvolveArgs fill_evolve_args(dstar_batch* b, const dstar_evolve_ctrl* ctrl, int trace_cap) {
    dstar_ctx* c = b->ctx;
    DsEvolveArgs a{};
    a.n_star = b->n_star; a.n_epoch = b->n_epoch; a.n_tab = DS_NTAB; a.zone_offset = b->zone_offset.p;
    a.tab_lnT = c->tab_lnT.p; a.tab_lnK = b->tab_lnK.p; a.tab_lnCp = b->tab_lnCp.p; a.tab_lnGamma = b->tab_lnGamma.p;
    a.tab_lnEnu = b->tab_lnEnu.p;
    a.cell_tab_zone0 = b->share[0].on ? b->share[0].d_tab0.p : nullptr;
    a.face_tab_zone0 = b->share[1].on ? b->share[1].d_tab0.p : nullptr;
    a.dm = b->dm.p; a.dm_bar = b->dm_bar.p; a.area = b->area.p; a.rho_bar = b->rho_bar.p; a.ePhi = b->ePhi.p;
    a.e2Phi = b->e2Phi.p; a.ePhi_bar = b->ePhi_bar.p; a.e2Phi_bar = b->e2Phi_bar.p; a.P = b->P.p;
    a.atm_index = b->atm_index.p; a.atm_nv = b->atm_nv; a.atm_lgTb = b->atm_lgTb.p; a.atm_lgTeff = b->atm_lgTeff.p;
    a.atm_lgflux = b->atm_lgflux.p; a.heat = b->heat.p; a.enuc_override = b->have_enuc ? b->enuc.p : nullptr;
    a.T_atm = b->T_atm.p; a.epoch_boundaries = b->epoch_boundaries.p; a.epoch_Mdots = b->epoch_Mdots.p;
    a.turn_on_extra_heating = ctrl->turn_on_extra_heating; a.fix_core_temperature = ctrl->fix_core_temperature;
    a.make_inner_boundary_insulating = ctrl->make_inner_boundary_insulating;
    a.fix_atmosphere_temperature_when_accreting = ctrl->fix_atmosphere_temperature_when_accreting;
    a.suppress_first_step_output = ctrl->suppress_first_step_output;
    a.starting_number_for_profile = ctrl->starting_number_for_profile;
    a.maximum_number_of_models = ctrl->maximum_number_of_models;
    a.integration_tolerance = ctrl->integration_tolerance; a.min_lg_T = ctrl->min_lg_temperature_integration;
    a.max_lg_T = ctrl->max_lg_temperature_integration; a.maximum_timestep = ctrl->maximum_timestep;
    a.lnT = b->lnT.p; a.dt = b->dt.p; a.tsec = b->tsec.p; a.model = b->model.p;
    a.t_monitor = b->t_mon.p; a.Teff_monitor = b->Teff_mon.p; a.Qb_monitor = b->Qb_mon.p; a.Teff = b->Teff.p;
    a.Lsurf = b->Lsurf.p; a.idid = b->idid.p; a.counters = b->counters.p;
    a.trace_cap = trace_cap; a.trace_count = trace_cap ? b->trace_count.p : nullptr; a.trace_model = b->trace_model.p;
    a.trace_tsec = b->tr_tsec.p; a.trace_dt = b->tr_dt.p; a.trace_Teff = b->tr_Teff.p; a.trace_Lsurf = b->tr_Lsurf.p;
    a.trace_Lnu = b->tr_Lnu.p; a.trace_Lnuc = b->tr_Lnuc.p; a.trace_Mdot = b->tr_Mdot.p; a.trace_ext = b->tr_ext.p;
    a.prof_interval = b->prof_interval; a.prof_cap = b->prof_cap; a.total_zones = b->total;
    a.prof_count = b->prof_cap ? b->prof_count.p : nullptr; a.prof_model = b->prof_model.p; a.prof_tsec = b->prof_tsec.p;
    a.prof_Mdot = b->prof_Mdot.p; a.prof_lnT = b->prof_lnT.p;
    return a;
}

int dstar_batch_set_evolve_mapping(dstar_batch* b, int mapping) {
    if (!b) return fail(DSTAR_ERR_ARG, "batch == NULL");
    if (mapping < 0 || mapping > 2) return fail(DSTAR_ERR_ARG, "evolve mapping: 0 automatic, 1 CTA per star, 2 warp per star");
    b->evolve_mapping = mapping;
    return DSTAR_OK;
}

int dstar_batch_evolve(dstar_batch* b, const dstar_evolve_ctrl* ctrl, int trace_cap) {
    if (!b || !ctrl) return fail(DSTAR_ERR_ARG, "batch_evolve");
    if (!b->have_evolve || !b->have_tables) return fail(DSTAR_ERR_STATE, "evolve inputs / tables missing");
    dstar_ctx* c = b->ctx;
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t ns = b->n_star;
    if (trace_cap < 0) trace_cap = 0;
    if (trace_cap != b->trace_cap || (trace_cap > 0 && !b->trace_model.p)) {
        const size_t tn = (size_t)trace_cap * ns;
        CK(b->trace_count.alloc(ns)); CK(b->trace_model.alloc(tn)); CK(b->tr_tsec.alloc(tn)); CK(b->tr_dt.alloc(tn));
        CK(b->tr_Teff.alloc(tn)); CK(b->tr_Lsurf.alloc(tn)); CK(b->tr_Lnu.alloc(tn)); CK(b->tr_Lnuc.alloc(tn));
        CK(b->tr_Mdot.alloc(tn)); CK(b->tr_ext.alloc(4 * tn));
        b->trace_cap = trace_cap;
    }
    const DsEvolveArgs a = fill_evolve_args(b, ctrl, trace_cap);
    // the star's rho_bar(1) is the fixed-up one (create_model.f:367-369): patch it into rho_bar
    // (the evolve kernel never reads face 0's rho_bar, L(1) comes from the atmosphere; kept for symmetry)
    CK(cudaStreamWaitEvent(s, b->ev_evolve_in, 0));   // evolve inputs were uploaded on the copy stream
    CK(cudaEventRecord(b->ev0, s));
    int launches = 0;
    // Mapping (dstar_batch_set_evolve_mapping).  Measured on B200 (profiles/r02_notes.md, r02d): the CTA-per-star form
    // is faster at every batch size tried (64 / 256 / 1024 stars of 253 zones: 2.2 / 2.6 / 9.6 ms against 6.6 / 10.2 /
    // 17.6 ms), so "automatic" is the CTA form; the warp form stays selectable.
    const bool warp_form = b->evolve_mapping == 2;
    if (!warp_form) {
        CK(launch_evolve_cta(a, nullptr, b->n_star, b->max_nz, c->prop.multiProcessorCount, b->evolve_minb, s));
        ++launches;
    }
    for (int q = 0; warp_form && q <= DS_N_EVOLVE_CLASS; ++q) {
        const dstar_batch::EvolveClass& ec = b->ev_class[q];
        if (ec.count == 0) continue;
        switch (ec.Z) {
            case 2: CK(launch_evolve_warp<2>(a, ec.list.p, ec.count, s)); break;
            case 4: CK(launch_evolve_warp<4>(a, ec.list.p, ec.count, s)); break;
            case 6: CK(launch_evolve_warp<6>(a, ec.list.p, ec.count, s)); break;
            case 8: CK(launch_evolve_warp<8>(a, ec.list.p, ec.count, s)); break;
            case 9: CK(launch_evolve_warp<9>(a, ec.list.p, ec.count, s)); break;
            case 10: CK(launch_evolve_warp<10>(a, ec.list.p, ec.count, s)); break;
            default:   // more than 320 zones: CTA-per-star form
                CK(launch_evolve_cta(a, ec.list.p, ec.count, ec.max_nz, c->prop.multiProcessorCount, b->evolve_minb, s));
        }
        CK(cudaGetLastError());
        ++launches;
    }
    CK(cudaEventRecord(b->ev1, s));
    CK(cudaEventSynchronize(b->ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1)); b->ms[2] = ms;
    CK(cudaEventRecord(b->ev_busy, s));
    b->launches = launches;
    return DSTAR_OK;
}

int dstar_batch_eval_rhs(dstar_batch* b, const dstar_evolve_ctrl* ctrl, int epoch, double* f, double* jac, double* surf) {
    if (!b || !ctrl) return fail(DSTAR_ERR_ARG, "batch_eval_rhs");
    if (!b->have_evolve || !b->have_tables) return fail(DSTAR_ERR_STATE, "evolve inputs / tables missing");
    if (epoch < 0 || epoch >= b->n_epoch) return fail(DSTAR_ERR_ARG, "epoch out of range");
    dstar_ctx* c = b->ctx;
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const DsEvolveArgs a = fill_evolve_args(b, ctrl, 0);
    DevBuf<double> d_f, d_jac, d_surf;
    CK(d_f.alloc(b->total)); CK(d_jac.alloc((size_t)3 * b->total)); CK(d_surf.alloc((size_t)3 * b->n_star));
    CK(cudaStreamWaitEvent(s, b->ev_evolve_in, 0));
    const int threads = ((b->max_nz + 31) / 32) * 32;
    if (threads <= 256) ds_evolve_rhs_test_kernel<256><<<b->n_star, threads, 0, s>>>(a, epoch, d_f.p, d_jac.p, d_surf.p);
    else ds_evolve_rhs_test_kernel<512><<<b->n_star, threads, 0, s>>>(a, epoch, d_f.p, d_jac.p, d_surf.p);
    CK(cudaGetLastError());
    if (f) CK(d_f.download(f, b->total, s));
    if (jac) CK(d_jac.download(jac, (size_t)3 * b->total, s));
    if (surf) CK(d_surf.download(surf, (size_t)3 * b->n_star, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_batch_set_profile_capture(dstar_batch* b, int interval, int capacity) {
    if (!b || interval < 0 || capacity < 0) return fail(DSTAR_ERR_ARG, "batch_set_profile_capture");
    CK(cudaSetDevice(b->ctx->device));
    if (interval == 0 || capacity == 0) { b->prof_interval = 0; b->prof_cap = 0; return DSTAR_OK; }
    const size_t ns = b->n_star;
    CK(b->prof_count.alloc(ns)); CK(b->prof_model.alloc((size_t)capacity * ns)); CK(b->prof_tsec.alloc((size_t)capacity * ns));
    CK(b->prof_Mdot.alloc((size_t)capacity * ns)); CK(b->prof_lnT.alloc((size_t)capacity * b->total));
    CK(cudaMemsetAsync(b->prof_count.p, 0, ns * sizeof(int), b->ctx->stream));
    b->prof_interval = interval; b->prof_cap = capacity;
    return DSTAR_OK;
}

int dstar_batch_download_profiles(dstar_batch* b, int32_t* count, int32_t* model, double* tsec, double* Mdot, double* lnT) {
    if (!b || b->prof_cap <= 0) return fail(DSTAR_ERR_STATE, "no profiles kept");
    if (!b->have_evolve) return fail(DSTAR_ERR_STATE, "nothing evolved");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t pn = (size_t)b->prof_cap * b->n_star;
    if (count) CK(b->prof_count.download(count, b->n_star, s));
    if (model) CK(b->prof_model.download(model, pn, s));
    if (tsec) CK(b->prof_tsec.download(tsec, pn, s));
    if (Mdot) CK(b->prof_Mdot.download(Mdot, pn, s));
    if (lnT) CK(b->prof_lnT.download(lnT, (size_t)b->prof_cap * b->total, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_batch_download_results(dstar_batch* b, double* t_monitor, double* Teff_monitor, double* Qb_monitor,
                                 int32_t* idid, int32_t* counters, double* lnT, double* Teff, double* Lsurf,
                                 double* dt, double* tsec, int32_t* model) {
    if (!b) return fail(DSTAR_ERR_ARG, "batch_download_results");
    if (!b->have_evolve) return fail(DSTAR_ERR_STATE, "nothing evolved");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t ns = b->n_star, ne = b->n_epoch;
    // The monitors are a few KB each: ten copies into pageable caller memory cost a driver round trip each.  They go
    // through one pinned staging block instead (async copies, one synchronisation, then plain memcpy to the caller).
    struct Item { void* dst; const void* src; size_t bytes; };
    const Item items[] = {
        {t_monitor, b->t_mon.p, ne * ns * sizeof(double)}, {Teff_monitor, b->Teff_mon.p, ne * ns * sizeof(double)},
        {Qb_monitor, b->Qb_mon.p, ne * ns * sizeof(double)}, {idid, b->idid.p, ne * ns * sizeof(int32_t)},
        {counters, b->counters.p, 7 * ns * sizeof(int32_t)}, {Teff, b->Teff.p, ns * sizeof(double)},
        {Lsurf, b->Lsurf.p, ns * sizeof(double)}, {dt, b->dt.p, ns * sizeof(double)},
        {tsec, b->tsec.p, ns * sizeof(double)}, {model, b->model.p, ns * sizeof(int32_t)}};
    size_t need = 0;
    for (const Item& it : items) if (it.dst) need += (it.bytes + 15) & ~(size_t)15;
    if (need > b->stage_bytes) {
        if (b->h_stage) cudaFreeHost(b->h_stage);
        b->h_stage = nullptr; b->stage_bytes = 0;
        CK(cudaMallocHost(&b->h_stage, need));
        b->stage_bytes = need;
    }
    size_t off = 0;
    for (const Item& it : items) {
        if (!it.dst) continue;
        CK(cudaMemcpyAsync(b->h_stage + off, it.src, it.bytes, cudaMemcpyDeviceToHost, s));
        off += (it.bytes + 15) & ~(size_t)15;
    }
    if (lnT) CK(b->lnT.download(lnT, b->total, s));
    CK(cudaStreamSynchronize(s));
    off = 0;
    for (const Item& it : items) {
        if (!it.dst) continue;
        std::memcpy(it.dst, b->h_stage + off, it.bytes);
        off += (it.bytes + 15) & ~(size_t)15;
    }
    return DSTAR_OK;
}

int dstar_batch_download_trace(dstar_batch* b, int32_t* count, int32_t* model, double* tsec, double* dt, double* Teff,
                               double* Lsurf, double* Lnu, double* Lnuc, double* Mdot) {
    if (!b || b->trace_cap <= 0) return fail(DSTAR_ERR_STATE, "no trace kept");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    const size_t tn = (size_t)b->trace_cap * b->n_star;
    if (count) CK(b->trace_count.download(count, b->n_star, s));
    if (model) CK(b->trace_model.download(model, tn, s));
    if (tsec) CK(b->tr_tsec.download(tsec, tn, s));
    if (dt) CK(b->tr_dt.download(dt, tn, s));
    if (Teff) CK(b->tr_Teff.download(Teff, tn, s));
    if (Lsurf) CK(b->tr_Lsurf.download(Lsurf, tn, s));
    if (Lnu) CK(b->tr_Lnu.download(Lnu, tn, s));
    if (Lnuc) CK(b->tr_Lnuc.download(Lnuc, tn, s));
    if (Mdot) CK(b->tr_Mdot.download(Mdot, tn, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_batch_download_trace_extrema(dstar_batch* b, double* ext) {
    if (!b || b->trace_cap <= 0) return fail(DSTAR_ERR_STATE, "no trace kept");
    if (!ext) return fail(DSTAR_ERR_ARG, "ext == NULL");
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t s = b->ctx->stream;
    CK(b->tr_ext.download(ext, (size_t)4 * b->trace_cap * b->n_star, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

double dstar_batch_last_kernel_ms(dstar_batch* b, int which) {
    if (!b || which < 0 || which > 3) return -1.0;
    return b->ms[which];
}
int dstar_batch_last_launch_count(dstar_batch* b) { return b ? b->launches : 0; }
int dstar_batch_last_redo_counts(dstar_batch* b, int32_t* counts) {
    if (!b || !counts) return fail(DSTAR_ERR_ARG, "batch_last_redo_counts");
    CK(cudaSetDevice(b->ctx->device));
    CK(b->redo_count.download(counts, 2, b->ctx->stream));
    CK(cudaStreamSynchronize(b->ctx->stream));
    if (micro_plan().split && micro_plan().blocks && b->blk_count.p) {   // parts scheduled by blocks: flagged 32-knot blocks
        std::vector<int32_t> bc(5 + 3 * DS_BLK_NKEY, 0);
        CK(b->blk_count.download(bc.data(), bc.size(), b->ctx->stream));
        CK(cudaStreamSynchronize(b->ctx->stream));
        if (micro_plan().blocks & 1) counts[0] = bc[3];
        if (micro_plan().blocks & 2) counts[1] = bc[4 + 3 * DS_BLK_NKEY];
    }
    return DSTAR_OK;
}

int dstar_micro_tables(dstar_ctx* ctx, int n_star, const int32_t* zone_offset, int ncharged, const double* rho,
                       const double* rho_bar, const double* P, const double* P_bar, const double* dm,
                       const double* ionic, const double* ionic_bar, const double* Yion, const double* Yion_bar,
                       const double* Zion, const double* Aion, const double* Tc_cell, const double* Tc_face,
                       const dstar_micro_ctrl* ctrl, double* tab_lnT, double* tab_lnCp, double* tab_lnGamma,
                       double* tab_lnEnu, double* tab_lnK, double* rho_bar1, int32_t* status) {
    dstar_batch* b = nullptr;
    int rc = dstar_batch_create(ctx, n_star, zone_offset, ncharged, &b);
    if (rc != DSTAR_OK) return rc;
    rc = dstar_batch_upload_structure(b, rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, Yion, Yion_bar, Zion, Aion,
                                      Tc_cell, Tc_face);
    if (rc == DSTAR_OK) rc = dstar_batch_micro_tables(b, ctrl, 3);
    if (rc == DSTAR_OK)
        rc = dstar_batch_download_tables(b, tab_lnT, tab_lnCp, tab_lnGamma, tab_lnEnu, tab_lnK, rho_bar1, status);
    dstar_batch_destroy(b);
    return rc;
}

int dstar_dsmath_eval(dstar_ctx* c, int fn, int n, const double* x, const double* y, double* out) {
    if (!c || n <= 0 || !x || !y || !out) return fail(DSTAR_ERR_ARG, "dsmath_eval");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf<double> dx, dy, dout;
    CK(dx.upload(x, n, s)); CK(dy.upload(y, n, s)); CK(dout.alloc(n));
    ds_dsmath_kernel<<<(n + 127) / 128, 128, 0, s>>>(fn, n, dx.p, dy.p, dout.p);
    CK(cudaGetLastError());
    CK(dout.download(out, n, s));
    CK(cudaStreamSynchronize(s));
    return DSTAR_OK;
}

int dstar_fp64_peak_tflops(dstar_ctx* c, double* tflops) {
    if (!c || !tflops) return fail(DSTAR_ERR_ARG, "fp64_peak");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int blocks = c->prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    DevBuf<double> out;
    CK(out.alloc((size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0, s));
        ds_fp64_peak_kernel<<<blocks, threads, 0, s>>>(out.p, iters, 1.0 + rep);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return DSTAR_OK;
}

}  // extern "C"
