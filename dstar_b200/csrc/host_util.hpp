// Host-side numerical utilities of the product (setup code above the C ABI, C++ because the
// reference's host language, Fortran, has no compiler in this image).  None of this is on the
// hot path: it builds the read-only tables the kernels consume.
#pragma once
#include <cmath>
#include <functional>
#include <string>
#include <vector>

namespace dsh_host {

// MESA interp_1d: piecewise-monotone cubic of Steffen (1990); f is (4,n), f(1,:) set on entry.
// Call sites replaced: load_sf.f:67, dStar_atm_mod.f:97-99, dStar_crust_mod.f:150-164.
void interp_pm(const double* x, int n, double* f);
// interp_value_and_slope / interp_value (call sites dStar_crust_lib.f:70-73,104; superfluid_lib.f:95)
void interp_value_and_slope(const double* x, int n, const double* f, double xv, double& val, double& slope);
double interp_value(const double* x, int n, const double* f, double xv);

// Hairer's DOP853 with dense output (the integrator behind MESA's dop853 used by
// NScool_crust_tov.f:173).  solout(nr, xold, x, y, dense) is called after every accepted step;
// dense(i, s) interpolates component i at xold <= s <= x.
struct Dop853 {
    using Rhs = std::function<void(double x, const double* y, double* dy)>;
    using Dense = std::function<double(int i, double s)>;
    using Solout = std::function<int(int nr, double xold, double x, const double* y, const Dense& dense)>;
    int n = 0;
    double rtol = 1e-6, atol = 1e-6;
    int max_steps = 100000;
    double max_step = 0.0;
    int nfcn = 0, nstep = 0, naccpt = 0, nrejct = 0;
    // returns idid (1 ok, 2 interrupted by solout, < 0 failure)
    int integrate(const Rhs& f, double& x, double* y, double xend, double h, const Solout& solout);
};

// safeguarded root finder for a monotone scalar function: bracket by expansion around x0, then
// Illinois-modified regula falsi to |dx| < epsx or |f| < epsy (stands in for MESA's
// look_for_brackets + safe_root_with_initial_guess, dStar_crust_mod.f:351-366)
double find_root(const std::function<double(double)>& f, double x0, double dx0, double epsx, double epsy,
                 int imax, int& ierr);

// Tc_data file (superfluid/private/load_sf.f:40-58): returns false on I/O failure
bool read_gap_file(const std::string& path, std::vector<double>& kF, std::vector<double>& Tc, double& kFmin,
                   double& kFmax);
// HELM binary written by tools/gen_helm_table.cpp: 21 arrays in helm_alloc.f:176-196 order
bool read_helm_file(const std::string& path, int& imax, int& jmax, std::vector<std::vector<double>>& arrays);

}  // namespace dsh_host
