// Internal seam between capi.cu and atm.cu (the bc09 kernels are a separate translation unit, see atm.cu).
#pragma once
#include <cuda_runtime.h>
struct DsHelmDev;
struct DsHostConsts;
cudaError_t ds_bc09_launch(const DsHelmDev& helm, const DsHostConsts& hc, int n_tab, int nv, const double* d_grav,
                           const double* d_Plight, const double* d_Pb, const double* d_lgTeff, double* d_lgTb,
                           double* d_info, int* d_ierr, cudaStream_t s);
