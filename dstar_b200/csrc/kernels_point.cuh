// Batched point evaluators behind the generic C-ABI entry points (eval_crust_eos, get_crust_neutrino_emissivity,
// get_thermal_conductivity, sf_get_results), the device-resident crust composition lookup
// (dStar_crust_get_composition_info, dStar_crust/public/dStar_crust_lib.f:109-147), and small helpers of the
// coefficient pass.  Non-template kernels: included by capi.cu only.
#pragma once
#include "kernels_micro.cuh"

// ---------------------------------------------------------------- batched point evaluators
struct DsEvalEosArgs {
    int n, ncharged;
    const double *rho, *T, *ionic, *Zion, *Aion, *Yion, *Tcs, *chi_in;
    const DsIonZ* ionz;
    DsEosCtrl eos;
    double* res;  // (15, n)
    double* comps;  // (7, 4, n) or null
    int* phase;
    double* chi_out;
    int* ierr;
};
__global__ void ds_eval_eos_kernel(DsEvalEosArgs a, DsHelmDev helm) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.n <= 0) return;
    const bool live = k < a.n;  // tail threads redo the last point and discard it (DS_PHASE needs all threads)
    if (!live) k = a.n - 1;
    DsComp c = ds_load_comp(a.ionic + (size_t)11 * k);
    double rho = a.rho[k], T = a.T[k];
    double chi = a.chi_in ? a.chi_in[k] : -1.0;
    if (chi == -1.0) chi = dse::nuclear_volume_fraction(rho, c, DS_DEFAULT_NUCLEAR_RADIUS);
    DsEosRes r;
    bool bad = false;
    const DsZonePre zp = ds_zone_pre(rho, c, chi);
    const DsHelmZone hz = ds_helm_zone(helm, rho, c.A / (1.0 - c.Yn), c.Z);
    double* cp = (a.comps && live) ? a.comps + (size_t)28 * k : nullptr;
    ds_eval_crust_eos<true>(helm, a.eos, rho, T, dsm::log10(T), c, zp, hz, a.ncharged, a.Zion, a.Aion, a.ionz,
                            a.Yion + (size_t)a.ncharged * k, a.Tcs[3 * (size_t)k + 1], chi, r, bad, cp);
    if (__syncthreads_or(bad)) {
        bool unused = false;
        ds_eval_crust_eos<false>(helm, a.eos, rho, T, dsm::log10(T), c, zp, hz, a.ncharged, a.Zion, a.Aion, a.ionz,
                                 a.Yion + (size_t)a.ncharged * k, a.Tcs[3 * (size_t)k + 1], chi, r, unused, cp);
    }
    if (!live) return;
    double* o = a.res + (size_t)15 * k;
    o[0] = r.lnP; o[1] = r.lnE; o[2] = r.lnS; o[3] = r.grad_ad; o[4] = r.chiRho; o[5] = r.chiT; o[6] = r.Cp;
    o[7] = r.Cv; o[8] = r.Chi; o[9] = r.Gamma; o[10] = r.Theta; o[11] = r.mu_e; o[12] = r.mu_n;
    o[13] = r.Gamma1; o[14] = r.Gamma3;
    a.phase[k] = r.phase;
    a.chi_out[k] = chi;
    a.ierr[k] = r.ierr;
}

struct DsEvalNuArgs {
    int n;
    const double *rho, *T, *ionic, *chi, *Tcn;
    int channels[5];
    double* eps;  // (6, n)
};
__global__ void ds_eval_nu_kernel(DsEvalNuArgs a) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.n <= 0) return;
    const bool live = k < a.n;
    if (!live) k = a.n - 1;
    DsNeutrino e;
    bool unused = false;
    const DsComp cc = ds_load_comp(a.ionic + (size_t)11 * k);
    dsn::ds_crust_neutrino<false>(a.rho[k], a.T[k], cc, ds_zone_pre(a.rho[k], cc, a.chi[k]), a.chi[k], a.Tcn[k],
                                  a.channels, e, unused);
    if (!live) return;
    double* o = a.eps + (size_t)6 * k;
    o[0] = e.total; o[1] = e.pair; o[2] = e.photo; o[3] = e.plasma; o[4] = e.brems; o[5] = e.pbf;
}

struct DsEvalCondArgs {
    int n;
    const double *rho, *T, *chi, *Gamma, *eta, *mu_e, *ionic, *Tns;
    DsCondCtrl cond;
    double* K;  // (10, n)
};
__global__ void ds_eval_cond_kernel(DsEvalCondArgs a) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.n <= 0) return;
    const bool live = k < a.n;
    if (!live) k = a.n - 1;
    DsCond c;
    const DsComp cc = ds_load_comp(a.ionic + (size_t)11 * k);
    dsk::ds_conductivity(a.cond, a.rho[k], a.T[k], a.chi[k], a.Gamma[k], a.eta[k], a.mu_e[k], cc,
                         dsk::ds_cond_pre(a.cond, a.rho[k], cc), a.Tns[k], nan(""), 0, c);
    if (!live) return;
    double* o = a.K + (size_t)10 * k;
    o[0] = c.total; o[1] = c.ee; o[2] = c.ei; o[3] = c.eQ; o[4] = c.sf; o[5] = c.nQ; o[6] = c.np; o[7] = c.kap;
    o[8] = c.electron_total; o[9] = c.neutron_total;
}

__global__ void ds_sf_kernel(int n, const double* kFp, const double* kFn, double* Tc, DsSfDev sf) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double t[3];
    ds_sf_get_results(sf, kFp[k], kFn[k], t);
    Tc[3 * (size_t)k] = t[0]; Tc[3 * (size_t)k + 1] = t[1]; Tc[3 * (size_t)k + 2] = t[2];
}

// ------------------------------------------------------------ crust composition lookup
// Device-resident composition table Y_i(lgP) (dStar_crust_def.f:10-24): per isotope a (4,nv)
// monotone-cubic coefficient block.  One thread per zone: interpolate every isotope, clip
// negatives, renormalise, moments with neutrons excluded (nucchem_lib.f:47-133).
struct DsCompArgs {
    int nz, nv, nisos, ncharged;
    const double* lgP_tab;  // (nv)
    const double* Ycoef;    // (4*nv, nisos)
    const double *Ziso, *Aiso;  // (nisos)
    // host-evaluated per-isotope powers: Z^(5/3), Z^(7/3), Z^2.5, Z*(Z+1)^1.5
    const double *Z53, *Z73, *Z52, *ZZ1;
    const double* lgP;  // (nz)
    double* ionic;      // (11, nz)
    double* Yion;       // (ncharged, nz)
};
__global__ void ds_crust_composition_kernel(DsCompArgs a) {
    int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= a.nz) return;
    double lgP_c = fmax(a.lgP[z], a.lgP_tab[0]);
    lgP_c = fmin(lgP_c, a.lgP_tab[a.nv - 1]);
    const int k = ds_locate(a.lgP_tab, a.nv, lgP_c);
    const double dx = lgP_c - a.lgP_tab[k];
    double* Yout = a.Yion + (size_t)a.ncharged * z;
    // pass 1: Xsum = sum(Y*A) in isotope order (dot_product, nucchem_lib.f:94)
    double Xsum = 0.0;
    int ic = 0;
    for (int i = 0; i < a.nisos; ++i) {
        const double* c = a.Ycoef + (size_t)4 * a.nv * i + 4 * k;
        double Y = c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
        if (Y < 0.0) Y = 0.0;
        Xsum += Y * a.Aiso[i];
    }
    // pass 2: renormalise, Ye, Yn in isotope order
    double Ye = 0.0, Yn = 0.0;
    ic = 0;
    for (int i = 0; i < a.nisos; ++i) {
        const double* c = a.Ycoef + (size_t)4 * a.nv * i + 4 * k;
        double Y = c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
        if (Y < 0.0) Y = 0.0;
        Y = Y / Xsum;
        Ye += Y * a.Ziso[i];
        if (a.Ziso[i] == 0.0 && a.Aiso[i] == 1.0) Yn += Y;
        if (a.Ziso[i] > 0.0 && a.Aiso[i] > 0.0) Yout[ic++] = Y;
    }
    if (Yn > 0.0 && Yn != 1.0)
        for (int j = 0; j < a.ncharged; ++j) Yout[j] = Yout[j] / (1.0 - Yn);
    double sA = 0, sZ = 0, sZ53 = 0, sZ2 = 0, sZ73 = 0, sZ52 = 0, sZZ1 = 0, sX = 0;
    ic = 0;
    for (int i = 0; i < a.nisos; ++i)
        if (a.Ziso[i] > 0.0 && a.Aiso[i] > 0.0) sA += fmin(1.0, fmax(Yout[ic++], 1.0e-50));
    const double A = 1.0 / sA;
    ic = 0;
    for (int i = 0; i < a.nisos; ++i) {
        if (!(a.Ziso[i] > 0.0 && a.Aiso[i] > 0.0)) continue;
        const double Y = Yout[ic++], Z = a.Ziso[i];
        sZ += Y * Z;
        sZ53 += Y * a.Z53[i];
        sZ2 += Y * (Z * Z);
        sZ73 += Y * a.Z73[i];
        sZ52 += Y * a.Z52[i];
        sZZ1 += Y * a.ZZ1[i];
        sX += Z * Z * Y / a.Aiso[i];
    }
    double* o = a.ionic + (size_t)11 * z;
    o[0] = A; o[1] = sZ * A; o[2] = sZ53 * A; o[3] = sZ2 * A; o[4] = sZ73 * A; o[5] = sZ52 * A;
    o[6] = sZZ1 * A; o[7] = sX; o[8] = Ye; o[9] = Yn; o[10] = o[3] - o[1] * o[1];
}

// table sharing: a star whose cell tables are another star's also takes that star's rho_bar(1) fix-up (create_model.f:367-369)
__global__ void ds_share_rho_bar1_kernel(int n_star, const int* __restrict__ owner, double* rho_bar1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_star && owner[i] != i) rho_bar1[i] = rho_bar1[owner[i]];
}

// ------------------------------------------------------------ FP64 peak probe (roofline denominator)
// Eight independent DFMA chains per thread; flops = 2 * 8 * iters * threads.
__global__ void ds_fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 0.999999, b = 1.0e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, m, b); a1 = __fma_rn(a1, m, b); a2 = __fma_rn(a2, m, b); a3 = __fma_rn(a3, m, b);
        a4 = __fma_rn(a4, m, b); a5 = __fma_rn(a5, m, b); a6 = __fma_rn(a6, m, b); a7 = __fma_rn(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
