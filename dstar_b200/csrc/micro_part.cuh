// Body of one coefficient-pass translation unit: the including .cu defines
//   DS_MICRO_PART   DS_PART_*            which part of the pass this unit holds
//   DS_MICRO_FN     ds_launch_micro_*    its launcher (micro_launch.hpp)
//   DS_MICRO_GS(X)  X(4) X(5) ...        the group counts to compile, the LAST one being the default
//   DS_MICRO_FAST   0/1                  also compile the branch-free (FAST = true) form
#include <algorithm>

#include "kernels_micro.cuh"
#include "micro_launch.hpp"

namespace {
template <int G, bool FAST>
cudaError_t launch_g(int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h, const DsSfDev& sf) {
    static_assert(16 + DS_PRE + DS_HZ + DS_CP <= DS_NTAB, "the prologue fetches the zone-level inputs with one lane each");
    constexpr int PART = DS_MICRO_PART;
    const int grid = std::max(1, std::min(sms * DS_MICRO_CTAS, (n_items + G - 1) / G));
    static_assert(DS_NTAB == 128 && DS_YMAX == 96, "ds_micro_dyn_bytes spells these out");
    const size_t dyn = ds_micro_dyn_bytes(PART, G);   // interpolation scratch + per-group abundances
    // static + dynamic shared memory exceed 48 KB: opt in, once per DEVICE (the attribute is per device and one process
    // may hold contexts on several)
    static bool once[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!once[dev & 63]) {
        if ((e = cudaFuncSetAttribute(ds_micro_kernel<PART, G, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)) != cudaSuccess) return e;
        once[dev & 63] = true;
    }
    ds_micro_kernel<PART, G, FAST><<<grid, DS_NTAB * G, dyn, s>>>(a, h, sf);
    return cudaGetLastError();
}
template <int G>
cudaError_t launch_f(bool fast, int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h, const DsSfDev& sf) {
#if DS_MICRO_FAST
    if (fast) return launch_g<G, true>(sms, n_items, s, a, h, sf);
#endif
    (void)fast;
    return launch_g<G, false>(sms, n_items, s, a, h, sf);
}
}  // namespace

cudaError_t DS_MICRO_FN(int G, bool fast, int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h,
                        const DsSfDev& sf, const DsHostConsts& hc) {
    static bool hc_set[64] = {};   // this unit's own copy of the host-evaluated constants, per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!hc_set[dev & 63]) {
        if ((e = cudaMemcpyToSymbol(ds_hc, &hc, sizeof hc)) != cudaSuccess) return e;
        hc_set[dev & 63] = true;
    }
    int last = 0;
#define DS_X(g) if (G == g) return launch_f<g>(fast, sms, n_items, s, a, h, sf); last = g;
    DS_MICRO_GS(DS_X)
#undef DS_X
    (void)last;
#define DS_X(g) if (last == g) return launch_f<g>(fast, sms, n_items, s, a, h, sf);
    DS_MICRO_GS(DS_X)
#undef DS_X
    return cudaErrorInvalidValue;
}

#ifdef DS_MICRO_BLOCK_FN
// The same part scheduled by 32-knot blocks in natural order (kernels_micro.cuh, ds_block_kernel): every warp works
// on its own (zone, knot block), no CTA barrier inside a round; then the interpolation of the part's table.
// a.blk_redo / a.blk_redo_count must be set, the count zeroed on the stream.
cudaError_t DS_MICRO_BLOCK_FN(int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h, const DsHostConsts& hc) {
    static bool hc_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!hc_set[dev & 63]) {
        if ((e = cudaMemcpyToSymbol(ds_hc, &hc, sizeof hc)) != cudaSuccess) return e;
        hc_set[dev & 63] = true;
    }
    if (n_items <= 0) return cudaSuccess;
    constexpr int PART = DS_MICRO_PART;
    const int grid = std::max(1, std::min(sms, (4 * n_items + DS_BLK_W - 1) / DS_BLK_W));
#if DS_MICRO_FAST
    ds_block_kernel<PART, true><<<grid, 32 * DS_BLK_W, 0, s>>>(a, h);
#endif
    ds_block_kernel<PART, false><<<grid, 32 * DS_BLK_W, 0, s>>>(a, h);
    ds_interp1_kernel<0><<<(n_items + DS_INTERP_G - 1) / DS_INTERP_G, DS_NTAB * DS_INTERP_G, 0, s>>>(a, PART == DS_PART_NU ? a.tab_lnEnu : a.tab_lnK);
    return cudaGetLastError();
}
#endif
