// Test-support kernel: evaluates one dsmath function on the device so the CPU-side copy of the
// specification can be compared bit for bit (tests/test_gpu_parity.py::test_dsmath_bits).
#pragma once
#include "dconst.cuh"

DSM_HD double ds_dsmath_dispatch(int fn, double x, double y) {
    switch (fn) {
        case 0: return dsm::exp(x);
        case 1: return dsm::log(x);
        case 2: return dsm::log10(x);
        case 3: return dsm::pow(x, y);
        case 4: return dsm::sin(x);
        case 5: return dsm::cos(x);
        case 6: return dsm::sinh(x);
        case 7: return dsm::cosh(x);
        default: return dsm::atan(x);
    }
}
__global__ void ds_dsmath_kernel(int fn, int n, const double* x, const double* y, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fn == 9) out[i] = x[i] / y[i];                       // IEEE division as the compiler emits it
    else if (fn == 10) {                                      // branch-free form (dconst.cuh) + its fallback
        bool bad = false;
        DS_DEN_SCOPE(true);
        const DsDen<true> d = den(y[i]);
        const double q = x[i] / d;
        out[i] = bad ? x[i] / DsDen<false>(y[i], nullptr) : q;
    } else if (fn == 11) {                                    // fast path only: NaN marks a raised flag
        bool bad = false;
        DS_DEN_SCOPE(true);
        const DsDen<true> d = den(y[i]);
        const double q = x[i] / d;
        out[i] = bad ? nan("") : q;
    }
    else if (fn == 12) out[i] = sqrt(x[i]);                   // IEEE square root as the compiler emits it
    else if (fn == 13) {                                      // branch-free form; NaN payload 0x7ff8..01 marks a raised flag
        bool bad = false;
        const double r = ds_sqrt<true>(x[i], &bad);
        out[i] = bad ? __longlong_as_double(0x7ff8000000000001ll) : r;
    }
    else out[i] = ds_dsmath_dispatch(fn, x[i], y[i]);
}
