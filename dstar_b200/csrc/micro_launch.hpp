// Launchers of the coefficient-pass kernels (kernels_micro.cuh).  Each part of the pass is its own translation unit
// of libdstar_b200.so (micro_eos.cu, micro_nu.cu, micro_face.cu, micro_cell.cu): the parts compile side by side, and
// each unit owns its copy of the __constant__ data (ds_hc is set by the launcher, once per device).
#pragma once
#include <cuda_runtime.h>

struct DsMicroArgs;
struct DsHelmDev;
struct DsSfDev;
struct DsHostConsts;

// One launch of ds_micro_kernel<PART, G, fast> on stream s: `sms` persistent CTAs at most, n_items work items (an upper
// bound for a redo pass, whose true count is read on the device).  G = zone groups (of 128 lanes) per CTA; a value the
// unit was not compiled for selects its default.  Returns the launch error.
cudaError_t ds_launch_micro_cell(int G, bool fast, int sms, int n_items, cudaStream_t s, const DsMicroArgs& a,
                                 const DsHelmDev& h, const DsSfDev& sf, const DsHostConsts& hc);   // EOS + neutrino fused
cudaError_t ds_launch_micro_eos(int G, bool fast, int sms, int n_items, cudaStream_t s, const DsMicroArgs& a,
                                const DsHelmDev& h, const DsSfDev& sf, const DsHostConsts& hc);    // lnCp, lnGamma, rho_bar(1)
// the EOS half scheduled by 32-knot blocks (classification, branch-free pass, plain pass, interpolation)
cudaError_t ds_launch_micro_eos_blocks(int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h,
                                       const DsHostConsts& hc);
// the neutrino half / the faces scheduled by 32-knot blocks in natural order (+ the interpolation of their table)
cudaError_t ds_launch_micro_nu_blocks(int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h,
                                      const DsHostConsts& hc);
cudaError_t ds_launch_micro_face_blocks(int sms, int n_items, cudaStream_t s, const DsMicroArgs& a, const DsHelmDev& h,
                                        const DsHostConsts& hc);
cudaError_t ds_launch_micro_nu(int G, bool fast, int sms, int n_items, cudaStream_t s, const DsMicroArgs& a,
                               const DsHelmDev& h, const DsSfDev& sf, const DsHostConsts& hc);     // lnEnu
cudaError_t ds_launch_micro_face(int G, bool fast, int sms, int n_items, cudaStream_t s, const DsMicroArgs& a,
                                 const DsHelmDev& h, const DsSfDev& sf, const DsHostConsts& hc);   // lnK
