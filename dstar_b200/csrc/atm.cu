// dStar_atm bc09 tables: second translation unit of libdstar_b200.so.  The evaluators (eos.cuh, conductivity.cuh)
// are compiled here WITHOUT the CTA-wide phase barriers the coefficient pass uses: the envelope integration is
// one thread per (table, knot) with data-dependent control flow (adaptive dop853, root finds).
#define DS_NO_PHASE_SYNC
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "atm_launch.hpp"
#include "kernels_atm.cuh"

namespace {
// compute_composition_moments for a two-species He4/Fe56 mix with neutrons excluded (nucchem_lib.f:47-133), as
// bc09.f:143-146 / :228-230 call it: Y = [1, 0]/A (He layer) or [0, 1]/A (Fe layer)
DsBc09Layer make_layer(int which, double ctf_fac) {
    const double Zs[2] = {2.0, 26.0}, As[2] = {4.0, 56.0};
    const double Y[2] = {(which == 0 ? 1.0 : 0.0) / As[0], (which == 0 ? 0.0 : 1.0) / As[1]};
    DsBc09Layer L{};
    double Ye = 0.0;
    for (int i = 0; i < 2; ++i) Ye += Y[i] * Zs[i];
    L.ncharged = 2;
    for (int i = 0; i < 2; ++i) { L.Zion[i] = Zs[i]; L.Aion[i] = As[i]; L.Yion[i] = Y[i]; L.ionz[i] = ds_ionz_compute(Zs[i], ctf_fac); }
    double sA = 0.0;
    for (int i = 0; i < 2; ++i) sA += std::min(1.0, std::max(Y[i], 1.0e-50));
    const double A = 1.0 / sA;
    double sZ = 0, sZ53 = 0, sZ2 = 0, sZ73 = 0, sZ52 = 0, sZZ1 = 0, sX = 0;
    for (int i = 0; i < 2; ++i) sZ += Y[i] * Zs[i];
    for (int i = 0; i < 2; ++i) sZ53 += Y[i] * dsm::pow(Zs[i], 5.0 / 3.0);
    for (int i = 0; i < 2; ++i) sZ2 += Y[i] * (Zs[i] * Zs[i]);
    for (int i = 0; i < 2; ++i) sZ73 += Y[i] * dsm::pow(Zs[i], 7.0 / 3.0);
    for (int i = 0; i < 2; ++i) sZ52 += Y[i] * dsm::pow(Zs[i], 2.5);
    for (int i = 0; i < 2; ++i) sZZ1 += Y[i] * (Zs[i] * dsm::pow(Zs[i] + 1.0, 1.5));
    for (int i = 0; i < 2; ++i) sX += Zs[i] * Zs[i] * Y[i] / As[i];
    L.c.A = A; L.c.Z = sZ * A; L.c.Z53 = sZ53 * A; L.c.Z2 = sZ2 * A; L.c.Z73 = sZ73 * A; L.c.Z52 = sZ52 * A;
    L.c.ZZ1_32 = sZZ1 * A; L.c.Z2XoA2 = sX; L.c.Ye = Ye; L.c.Yn = 0.0; L.c.Q = L.c.Z2 - L.c.Z * L.c.Z;
    return L;
}
}  // namespace

cudaError_t ds_bc09_launch(const DsHelmDev& helm, const DsHostConsts& hc, int n_tab, int nv, const double* d_grav,
                           const double* d_Plight, const double* d_Pb, const double* d_lgTeff, double* d_lgTb,
                           double* d_info, int* d_ierr, cudaStream_t s) {
    cudaError_t e = cudaMemcpyToSymbolAsync(ds_hc, &hc, sizeof hc, 0, cudaMemcpyHostToDevice, s);   // this unit's own copy
    if (e != cudaSuccess) return e;
    DsBc09Args a{};
    a.n_tab = n_tab; a.nv = nv; a.grav = d_grav; a.Plight = d_Plight; a.Pb = d_Pb; a.lgTeff = d_lgTeff; a.lgTb = d_lgTb;
    a.info = d_info; a.ierr = d_ierr;
    a.layer[0] = make_layer(0, hc.ctf_fac);
    a.layer[1] = make_layer(1, hc.ctf_fac);
    a.eos = DsEosCtrl{175.0, 140.0, DS_R4(1.0e-9)};       // a fresh eos handle: dStar_eos_def.f:58-62
    a.cond = DsCondCtrl{0, 0, 1, 1, 10.0, 9.0};           // bc09.f:69-72
    a.max_iter_photosphere = 200; a.max_iter_density = 60;
    a.tol_photosphere_lnrho = 1.0e-13; a.tol_photosphere_condition = 0.0; a.tol_density = 1.0e-13;
    // the envelope integration nests root finds inside dop853 inside the layer loop around a ~20 KB-frame evaluator
    const int total = n_tab * nv;
    ds_bc09_kernel<<<(total + 31) / 32, 32, 0, s>>>(a, helm);
    return cudaGetLastError();
}
