// Coefficient pass, cell half 1: eval_crust_eos per (zone, knot) -> lnCp, lnGamma tables and the rho_bar(1) fix-up
// (NScool_create_model.f:344-369, :386-397).
// alignment barriers between the components only, not inside ee_exchange (4.04 -> 4.00 ms).
#ifndef DS_EOS_PHASE_INNER  // experiment switch: keep the eight barriers inside ee_exchange
#define DS_NO_PHASE_INNER
#endif
#ifdef DS_EOS_NO_PHASE     // experiment: no alignment barrier at all in the EOS unit
#define DS_NO_PHASE_SYNC
#endif
#ifdef DS_EOS_NO_PHASE_ION // experiment: none inside the species loop of ion_mixture
#define DS_NO_PHASE_ION
#endif
// one-test validity check of the branch-free divisions (dconst.cuh, DS_DEN_LEAN): the EOS has no quotients below 2^-650
// (no block of any workload takes the plain-division pass); EOS half 3.45 -> 3.41 ms, INT-16 10.35 -> 10.09.
// -DDS_DEN_FULL_EOS keeps the three-test form.
#ifndef DS_DEN_FULL_EOS
#define DS_DEN_LEAN
#endif
#define DS_MICRO_PART DS_PART_EOS
#define DS_MICRO_FN ds_launch_micro_eos
#ifdef DS_MICRO_SCAN
#define DS_MICRO_GS(X) X(3) X(4) X(5) X(6) X(8) X(7)
#else
#define DS_MICRO_GS(X) X(7)
#endif
#define DS_MICRO_FAST 1
#include "micro_part.cuh"

// The same half scheduled by 32-knot blocks (kernels_micro.cuh, ds_block_kernel): counting sort of the blocks by
// phase class and species count, branch-free pass over the three class lists, plain pass over the flagged blocks,
// interpolation.  a.blk_count must be zeroed on the stream before the call.  Seven launches.
cudaError_t ds_launch_micro_eos_blocks(int sms, int n_items, cudaStream_t s, const DsMicroArgs& a_in, const DsHelmDev& h,
                                       const DsHostConsts& hc) {
    static bool hc_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!hc_set[dev & 63]) {
        if ((e = cudaMemcpyToSymbol(ds_hc, &hc, sizeof hc)) != cudaSuccess) return e;
        hc_set[dev & 63] = true;
    }
    if (n_items <= 0) return cudaSuccess;
    DsMicroArgs a = a_in;
    a.blk_redo = a.blk_list + (size_t)3 * 4 * a.total_zones; a.blk_redo_count = a.blk_count + 3;
    const size_t dyn = (size_t)DS_BLK_W * DS_YMAX * sizeof(double);
    const int grid = std::max(1, std::min(sms, (4 * n_items + DS_BLK_W - 1) / DS_BLK_W));
    ds_eos_classify_kernel<0><<<(n_items + 127) / 128, 128, 0, s>>>(a);   // histogram of the sort keys
    ds_eos_prefix_kernel<0><<<1, 32, 0, s>>>(a.blk_count);
    ds_eos_classify_kernel<1><<<(n_items + 127) / 128, 128, 0, s>>>(a);   // scatter
    ds_block_kernel<DS_PART_EOS, true><<<grid, 32 * DS_BLK_W, dyn, s>>>(a, h);
    ds_block_kernel<DS_PART_EOS, false><<<grid, 32 * DS_BLK_W, dyn, s>>>(a, h);
    ds_eos_interp_kernel<0><<<(n_items + DS_INTERP_G - 1) / DS_INTERP_G, DS_NTAB * DS_INTERP_G, 0, s>>>(a);
    return cudaGetLastError();
}
