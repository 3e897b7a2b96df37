// dStar_atm 'bc09' envelope tables on the device -- the reference's DEFAULT outer boundary condition
// (NScool/defaults/controls.defaults:149; dStar_atm/private/bc09.f:45-666).
//
// One table = 256 T_eff knots (dStar_atm_def.f:4-8); per knot the reference finds the photosphere
// (P = tau g / kappa, bc09.f:284-443), then integrates d ln T^4 / d ln P through a He and an Fe layer with
// dop853 (bc09.f:97-282), the right-hand side inverting the EOS for rho(P, T) by a root find (bc09.f:553-666) and
// evaluating EOS + conductivity for the opacity and the adiabatic cap (bc09.f:445-551).  Per knot that is ~5000
// dependent EOS evaluations; knots and tables are independent, so the mapping is ONE THREAD PER (table, knot):
// a table is 8 warps, a sweep over (g, P_light, P_b) -- examples/wrapper/src/run.f:89-95 -- fills the machine.
// The evaluators are the same inlined device functions the coefficient pass uses (plain IEEE divisions, phase
// barriers compiled out in this translation unit because control flow here is per thread).
//
// Departure from the reference, on purpose: the reference walks the knots from hot to cold and hands the
// photospheric density of one knot to the next as the root finder's first guess (bc09.f:77-82), which would
// serialise a table.  Here every knot starts from its own ideal-gas + Thomson guess and all roots are converged to
// 1e-13, i.e. well inside the tolerances the reference itself asks of MESA's root finder (1e-6 in ln rho_ph, 1e-12
// in ln rho); the tests' CPU checker implements exactly this form too and is compared knot by knot.
#pragma once
#include "conductivity.cuh"
#include "dconst.cuh"
#include "eos.cuh"

struct DsBc09Layer {   // He (0) or Fe (1): moments with neutrons excluded, as bc09.f:143-177 / :228-253 build them
    DsComp c;
    int ncharged;
    double Zion[2], Aion[2], Yion[2];
    DsIonZ ionz[2];
};
struct DsBc09Args {
    int n_tab, nv;
    const double *grav, *Plight, *Pb;  // (n_tab)
    const double* lgTeff;              // (nv) ascending knots
    double* lgTb;                      // (nv, n_tab)
    double* info;                      // (4, nv, n_tab) rho_ph, P_ph, kappa_ph, number of EOS evaluations -- or null
    int* ierr;                         // (nv, n_tab)
    DsBc09Layer layer[2];
    DsEosCtrl eos;
    DsCondCtrl cond;
    int max_iter_photosphere, max_iter_density;
    double tol_photosphere_lnrho, tol_photosphere_condition, tol_density;
};

namespace dsa {
using namespace dsc;

struct Par {   // rpar / ipar of bc09.f:6-34
    const DsHelmDev* helm;
    const DsBc09Args* a;
    const DsBc09Layer* L;
    double grav, tau, Teff, pres, temp, rho, kappa;
    int n_eos;
};
struct EosOut {
    double lnP, chiRho, grad_ad, Gamma, Theta, mu_e, chi;
};

// eval_crust_eos (dStar_eos_lib.f:123) at one point; never inlined so the ~0.4 MB evaluator exists once in the kernel
__device__ __noinline__ void eos_point(const Par& p, double rho, double T, EosOut& o) {
    const DsBc09Layer& L = *p.L;
    const double chi = dse::nuclear_volume_fraction(rho, L.c, DS_DEFAULT_NUCLEAR_RADIUS);
    const DsZonePre zp = ds_zone_pre(rho, L.c, chi);
    const DsHelmZone hz = ds_helm_zone(*p.helm, rho, L.c.A / (1.0 - L.c.Yn), L.c.Z);
    DsEosRes r;
    bool bad = false;
    ds_eval_crust_eos<false>(*p.helm, p.a->eos, rho, T, dsm::log10(T), L.c, zp, hz, L.ncharged, L.Zion, L.Aion, L.ionz,
                             L.Yion, 0.0, chi, r, bad);
    o.lnP = r.lnP; o.chiRho = r.chiRho; o.grad_ad = r.grad_ad; o.Gamma = r.Gamma; o.Theta = r.Theta; o.mu_e = r.mu_e;
    o.chi = chi;
}
// kappa = 4/3 a c T^3 / rho / K_total (bc09.f:437, :549)
__device__ __noinline__ double opacity_point(const Par& p, double rho, double T, const EosOut& e) {
    DsCond K;
    dsk::ds_conductivity(p.a->cond, rho, T, e.chi, e.Gamma, e.Theta, e.mu_e, p.L->c, dsk::ds_cond_pre(p.a->cond, rho, p.L->c),
                         0.0, nan(""), 0, K);
    return 4.0 * onethird * arad * clight * dsd::ipow<3>(T) / rho / K.total;
}

// WHICH = 0: photosphere (bc09.f:373-443), 1: eval_pressure (bc09.f:607-666)
template <int WHICH>
DS_D double root_fn(Par& p, double lnrho, double& dfdx, int& ierr) {
    ierr = 0;
    const double rho = dsm::exp(lnrho);
    EosOut e;
    ++p.n_eos;
    if (WHICH == 0) {
        eos_point(p, rho, p.Teff, e);
        const double P = dsm::exp(e.lnP);
        const double kappa = opacity_point(p, rho, p.Teff, e);
        p.kappa = kappa;
        p.pres = P;
        dfdx = 0.0;
        return P - p.tau * p.grav / kappa;
    }
    eos_point(p, rho, p.temp, e);
    if (!isfinite(e.lnP)) { ierr = -9; return 0.0; }
    dfdx = e.chiRho;
    return e.lnP - dsm::log(p.pres);
}

// MESA num_lib look_for_brackets / safe_root_with_initial_guess, restated from their documented contract: bracket
// search by a doubling increment around the guess; Newton step when a derivative is supplied and it stays inside the
// bracket, bisection otherwise; stop on |dx| <= epsx or |f| <= epsy
template <int WHICH>
DS_D int look_for_brackets(Par& p, double x, double dx, double& x1, double& x3, double& y1, double& y3, int imax) {
    int ierr = 0;
    double dfdx;
    x1 = x - dx;
    x3 = x + dx;
    y1 = root_fn<WHICH>(p, x1, dfdx, ierr);
    if (ierr != 0) return ierr;
    y3 = root_fn<WHICH>(p, x3, dfdx, ierr);
    if (ierr != 0) return ierr;
    for (int i = 0; i < imax; ++i) {
        if (!(y1 * y3 > 0.0)) return (y1 * y3 <= 0.0) ? 0 : -1;
        dx = 2.0 * dx;
        if (fabs(y1) < fabs(y3)) {
            x3 = x1; y3 = y1;
            x1 = x1 - dx;
            y1 = root_fn<WHICH>(p, x1, dfdx, ierr);
        } else {
            x1 = x3; y1 = y3;
            x3 = x3 + dx;
            y3 = root_fn<WHICH>(p, x3, dfdx, ierr);
        }
        if (ierr != 0) return ierr;
    }
    return (y1 * y3 <= 0.0) ? 0 : -1;
}
template <int WHICH>
DS_D double safe_root(Par& p, double x_guess, double x1, double x3, double y1, double y3, int imax, double epsx,
                      double epsy, int& ierr) {
    ierr = 0;
    if (y1 == 0.0) return x1;
    if (y3 == 0.0) return x3;
    double x = x_guess;
    if (!(x > x1 && x < x3)) x = 0.5 * (x1 + x3);
    for (int i = 0; i < imax; ++i) {
        double dfdx = 0.0;
        const double y = root_fn<WHICH>(p, x, dfdx, ierr);
        if (ierr != 0) return x;
        if (fabs(y) <= epsy) return x;
        if ((y > 0.0) == (y1 > 0.0)) { x1 = x; y1 = y; } else { x3 = x; y3 = y; }
        double xn = 0.5 * (x1 + x3);
        if (dfdx != 0.0) {
            const double xt = x - y / dfdx;
            if (xt > x1 && xt < x3) xn = xt;
        }
        if (fabs(xn - x) <= epsx) return xn;
        x = xn;
    }
    ierr = -1;
    return x;
}

// deriv + get_coefficients + get_lnrho_from_PT (bc09.f:445-605); returns ierr
DS_D int deriv(Par& p, double lnP, double lnT4, double& dlnT4dlnP) {
    double rho = p.rho;
    const double P = dsm::exp(lnP);
    const double T = p.Teff * dsm::exp(0.25 * lnT4);
    // get_lnrho_from_PT
    p.temp = T;
    p.pres = P;
    const double lnrho_guess = dsm::log(rho);
    double x1, x3, y1, y3;
    int ierr = look_for_brackets<1>(p, lnrho_guess, 0.5, x1, x3, y1, y3, p.a->max_iter_density);
    if (ierr != 0) return ierr;
    const double lnrho = safe_root<1>(p, lnrho_guess, x1, x3, y1, y3, p.a->max_iter_density, p.a->tol_density,
                                      p.a->tol_density, ierr);
    if (ierr != 0) return ierr;
    rho = dsm::exp(lnrho);
    EosOut e;
    ++p.n_eos;
    eos_point(p, rho, T, e);
    if (!(fabs(dsm::exp(e.lnP) - P) < 1.0e-3 * P)) return -10;   // assertion "EOS is behaving", bc09.f:536
    const double kappa = opacity_point(p, rho, T, e);
    const double v = 0.75 * kappa * P / p.grav / dsm::exp(lnT4);
    dlnT4dlnP = fmin(v, 4.0 * e.grad_ad);
    p.rho = rho;   // the new anchor point (bc09.f:481)
    return 0;
}

// Scalar DOP853 (Hairer, Norsett & Wanner; MESA num_lib dop853 as bc09.f:202,224 calls it): one equation, scalar
// tolerances, given first step, no dense output.  A right-hand side that fails (ierr /= 0: "retry with smaller
// timestep", bc09.f:458 -- e.g. a trial stage at a temperature whose radiation pressure alone exceeds P, where no
// density solves the EOS) rejects the step with the controller's largest reduction, h / facc1.
DS_D int dop853_scalar(Par& p, double& x, double& y, double xend, double h0, int max_steps, double rtol, double atol) {
    const double uround = 2.3e-16, safe = 0.9, fac1 = 0.333, fac2 = 6.0;
    const double facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const double hmax = fabs(xend - x);
    const double posneg = copysign(1.0, xend - x);
    double k[13];
    double facold = 1.0e-4;
    bool last = false, reject = false;
    int nstep = 0;
    if (deriv(p, x, y, k[0]) != 0) return -5;
    double h = h0;
    int idid = 1;
    for (;;) {
        if (nstep > max_steps) { idid = -2; break; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { idid = -3; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
        ++nstep;
        bool failed = false;
#pragma unroll 1
        for (int s = 1; s < 12 && !failed; ++s) {
            double acc = 0.0;
            for (int j = 0; j < s; ++j) {
                const double a = dop_a[s][j];
                if (a != 0.0) acc += a * k[j];
            }
            failed = deriv(p, x + dop_c[s] * h, y + h * acc, k[s]) != 0;
        }
        if (failed) {
            h = h / facc1;
            reject = true;
            last = false;
            continue;
        }
        const double xph = x + h;
        double k4 = 0.0;
        for (int j = 0; j < 12; ++j)
            if (dop_b[j] != 0.0) k4 += dop_b[j] * k[j];
        const double ynew = y + h * k4;
        const double sk = atol + rtol * fmax(fabs(y), fabs(ynew));
        double e3 = 0.0, e5 = 0.0;
        for (int j = 0; j < 12; ++j) {
            if (dop_e3[j] != 0.0) e3 += dop_e3[j] * k[j];
            if (dop_e5[j] != 0.0) e5 += dop_e5[j] * k[j];
        }
        const double err2 = (e3 / sk) * (e3 / sk);
        double err = (e5 / sk) * (e5 / sk);
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (1 * deno));
        const double fac11 = dsm::pow(err, 1.0 / 8.0);
        double fac = fac11 / dsm::pow(facold, 0.0);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0 && deriv(p, xph, ynew, k[12]) != 0) {   // the new point itself cannot be evaluated
            h = h / facc1;
            reject = true;
            last = false;
            continue;
        }
        if (err <= 1.0) {
            facold = fmax(err, 1.0e-4);
            k[0] = k[12];
            y = ynew;
            x = xph;
            if (last) { idid = 1; break; }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = false;
        } else {
            hnew = h / fmin(facc1, fac11 / safe);
            reject = true;
            last = false;
        }
        h = hnew;
    }
    return idid;
}

// find_photospheric_pressure (bc09.f:284-371) with the ideal-gas + Thomson first guess (:322-333)
DS_D int find_photospheric_pressure(Par& p, double& rho_ph, double& P_ph, double& kappa) {
    const double Teff = p.Teff;
    const double fallback_Pphoto = 2.0 * onethird * arad * dsd::ipow<4>(Teff);
    const double kappa_Th = Thomson * avogadro * p.L->c.Ye;
    const double eps_ph = p.a->tol_photosphere_condition * p.tau * p.grav / kappa_Th;
    const double Pgas = p.tau * p.grav / kappa_Th - onethird * arad * dsd::ipow<4>(Teff);
    if (Pgas < 0.0) { P_ph = fallback_Pphoto; kappa = kappa_Th; return -2; }
    const double lnrho_guess = dsm::log(Pgas * amu * p.L->c.A / (p.L->c.Z + 1.0) / boltzmann / Teff);
    double x1, x3, y1, y3;
    int ierr = look_for_brackets<0>(p, lnrho_guess, 0.1, x1, x3, y1, y3, p.a->max_iter_photosphere);
    if (ierr != 0) { P_ph = fallback_Pphoto; rho_ph = dsm::exp(lnrho_guess); kappa = p.kappa; return ierr; }
    const double lnrho_ph = safe_root<0>(p, lnrho_guess, x1, x3, y1, y3, p.a->max_iter_photosphere,
                                         p.a->tol_photosphere_lnrho, eps_ph, ierr);
    if (ierr != 0) { P_ph = fallback_Pphoto; rho_ph = dsm::exp(lnrho_guess); kappa = p.kappa; return ierr; }
    double dummy;
    root_fn<0>(p, lnrho_ph, dummy, ierr);
    P_ph = p.pres;
    rho_ph = dsm::exp(lnrho_ph);
    kappa = p.kappa;
    return 0;
}

// do_integrate_bc09_atm (bc09.f:97-282)
DS_D int integrate_bc09_atm(Par& p, double lgyb, double lgy_light, double lgTeff, double& lgTb, double* info) {
    p.Teff = dsd::dexp10(lgTeff);
    p.tau = twothird;
    p.L = &p.a->layer[0];
    double rho = -1.0, P = 0.0, kappa = 0.0;
    int ierr = find_photospheric_pressure(p, rho, P, kappa);
    if (info) { info[0] = rho; info[1] = P; info[2] = kappa; }
    if (ierr != 0) return ierr;
    p.rho = rho; p.temp = p.Teff; p.pres = P; p.kappa = kappa;
    double lnP = dsm::log(P);
    double lnT4 = 0.0;
    const double lg_grav = dsm::log10(p.grav);
    double lnPend = fmax(ln10 * (lgy_light + lg_grav), lnP + ln10);
    lnPend = fmin(lnPend, ln10 * (lgyb + lg_grav));
    int idid = dop853_scalar(p, lnP, lnT4, lnPend, 0.001, 10000, 1.0e-6, 1.0e-6);
    if (idid < 0) return idid;
    p.L = &p.a->layer[1];
    lnPend = ln10 * (lgyb + lg_grav);
    if (lnP < lnPend) {
        idid = dop853_scalar(p, lnP, lnT4, lnPend, 0.001, 10000, 1.0e-6, 1.0e-6);
        if (idid < 0) return idid;
    }
    lgTb = 0.25 * lnT4 / ln10 + lgTeff;
    return 0;
}
}  // namespace dsa

// thread = (table, knot); 32-thread CTAs so that the 8 warps of one table land on 8 different SMs
__global__ void __launch_bounds__(32) ds_bc09_kernel(DsBc09Args a, DsHelmDev helm) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= a.n_tab * a.nv) return;
    const int t = gid / a.nv, i = gid % a.nv;
    dsa::Par p;
    p.helm = &helm;
    p.a = &a;
    p.grav = a.grav[t];
    p.n_eos = 0;
    const double lgyb = dsm::log10(a.Pb[t] / p.grav);
    const double lgy_light = dsm::log10(a.Plight[t] / p.grav);
    double lgTb = 0.0;
    double* info = a.info ? a.info + (size_t)4 * gid : nullptr;
    const int ierr = dsa::integrate_bc09_atm(p, lgyb, lgy_light, a.lgTeff[i], lgTb, info);
    if (info) info[3] = (double)p.n_eos;
    a.lgTb[gid] = lgTb;
    a.ierr[gid] = ierr;
}
