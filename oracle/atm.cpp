// ORACLE (test infrastructure): the bc09 envelope model of dStar_atm -- the reference's DEFAULT outer
// boundary (NScool/defaults/controls.defaults:149) -- restated from
//   dStar_atm/private/bc09.f:45-95    do_get_bc09_Teff        (knot loop, hottest first, rho_ph carried down)
//   dStar_atm/private/bc09.f:97-282   do_integrate_bc09_atm   (He layer, then Fe layer, dop853 in ln P)
//   dStar_atm/private/bc09.f:284-371  find_photospheric_pressure
//   dStar_atm/private/bc09.f:373-443  photosphere             (P - tau g / kappa)
//   dStar_atm/private/bc09.f:445-483  deriv
//   dStar_atm/private/bc09.f:485-551  get_coefficients
//   dStar_atm/private/bc09.f:553-605  get_lnrho_from_PT
//   dStar_atm/private/bc09.f:607-666  eval_pressure
//   dStar_atm/private/dStar_atm_mod.f:46-106  do_generate_atm_table (256 knots in lgTeff, interp_pm in lgTb)
//
// Third-party arithmetic not in the tree (MESA r15140 num_lib): look_for_brackets,
// safe_root_with_initial_guess.  Their contract is restated from the call sites (bracket search by a growing
// increment around the guess; "safe" root = Newton accelerated by the derivative when one is supplied, bisection
// otherwise, inside the bracket, stopped by epsx on x or epsy on |f|, at most imax function calls).  The iterates
// of MESA's own implementation are not reproducible without its source, so a root is pinned only to the
// tolerances the reference passes: 1e-12 for the density inversion (bc09.f:568), 1e-6 in ln(rho) for the
// photosphere (bc09.f:291-293).  Bc09Settings::reference() uses exactly those; Bc09Settings::tight() converges
// every root to 1e-13 and starts each knot from its own ideal-gas guess, which makes knots independent (the
// device kernel's form).  tests/test_oracle_atm.py bounds the difference between the two.
// Pinned by the bc09 blocks of dStar_atm/test/sample_output.data (ndiff relerr 1e-3, dStar_atm/test/chk:15).
#include <algorithm>
#include <cmath>

#include "dstar_oracle.hpp"

namespace dso {

namespace {

typedef double (*RootFn)(double x, double& dfdx, void* ctx, int& ierr);

// MESA num_lib look_for_brackets(x, dx, x1, x3, f, y1, y3, imax, ...): ierr = 0 when y1*y3 <= 0
int look_for_brackets(double x, double dx, double& x1, double& x3, RootFn f, void* ctx, double& y1, double& y3,
                      int imax) {
    int ierr = 0;
    double dfdx;
    x1 = x - dx;
    x3 = x + dx;
    y1 = f(x1, dfdx, ctx, ierr);
    if (ierr != 0) return ierr;
    y3 = f(x3, dfdx, ctx, ierr);
    if (ierr != 0) return ierr;
    for (int i = 0; i < imax; ++i) {
        if (!(y1 * y3 > 0.0)) return (y1 * y3 <= 0.0) ? 0 : -1;
        dx = 2.0 * dx;
        if (std::fabs(y1) < std::fabs(y3)) {   // the root lies beyond x1: slide the window down
            x3 = x1; y3 = y1;
            x1 = x1 - dx;
            y1 = f(x1, dfdx, ctx, ierr);
        } else {
            x1 = x3; y1 = y3;
            x3 = x3 + dx;
            y3 = f(x3, dfdx, ctx, ierr);
        }
        if (ierr != 0) return ierr;
    }
    return (y1 * y3 <= 0.0) ? 0 : -1;
}

// MESA num_lib safe_root_with_initial_guess(f, x_guess, x1, x3, y1, y3, imax, epsx, epsy, ...)
double safe_root_with_initial_guess(RootFn f, void* ctx, double x_guess, double x1, double x3, double y1, double y3,
                                    int imax, double epsx, double epsy, int& ierr) {
    ierr = 0;
    if (y1 == 0.0) return x1;
    if (y3 == 0.0) return x3;
    double x = x_guess;
    if (!(x > x1 && x < x3)) x = 0.5 * (x1 + x3);
    for (int i = 0; i < imax; ++i) {
        double dfdx = 0.0;
        const double y = f(x, dfdx, ctx, ierr);
        if (ierr != 0) return x;
        if (std::fabs(y) <= epsy) return x;
        if ((y > 0.0) == (y1 > 0.0)) { x1 = x; y1 = y; } else { x3 = x; y3 = y; }
        double xn = 0.5 * (x1 + x3);
        if (dfdx != 0.0) {
            const double xt = x - y / dfdx;
            if (xt > x1 && xt < x3) xn = xt;
        }
        if (std::fabs(xn - x) <= epsx) return xn;
        x = xn;
    }
    ierr = -1;
    return x;
}

struct Par {   // rpar / ipar of bc09.f:6-34
    const HelmTable* h;
    EosControls eos;
    CondControls cond;
    Composition ionic;
    int ncharged;
    double Zion[2], Aion[2], Yion[2];
    double grav, tau, Teff, pres, temp, rho, kappa;
    const Bc09Settings* st;
    long n_eos = 0;
};

void set_layer(Par& p, int which) {   // bc09.f:143-177 (He) and :228-253 (Fe)
    const double Zs[2] = {2.0, 26.0}, As[2] = {4.0, 56.0};
    double Y[2] = {(which == 0 ? 1.0 : 0.0) / As[0], (which == 0 ? 0.0 : 1.0) / As[1]};
    MomentsOptions o;
    o.exclude_neutrons = true;
    double Xsum;
    int idx[2];
    compute_composition_moments(2, Zs, As, Y, o, p.ionic, Xsum, p.ncharged, idx, p.Yion);
    for (int i = 0; i < p.ncharged; ++i) { p.Zion[i] = Zs[idx[i]]; p.Aion[i] = As[idx[i]]; }
}

// bc09.f:373-443
double photosphere(double lnrho, double& dfdlnrho, void* ctx, int& ierr) {
    Par& p = *static_cast<Par*>(ctx);
    ierr = 0;
    const double rho = dsm::exp(lnrho);
    const double Tcs[3] = {0.0, 0.0, 0.0};
    double res[num_eos_results], chi = use_default_nuclear_size;
    int phase;
    eval_crust_eos(*p.h, p.eos, rho, p.Teff, p.ionic, p.ncharged, p.Zion, p.Aion, p.Yion, Tcs, res, phase, chi);
    ++p.n_eos;
    const double P = dsm::exp(res[i_lnP]);
    CondComponents K;
    conductivity(p.cond, rho, p.Teff, chi, res[i_Gamma], res[i_Theta], res[i_mu_e], p.ionic, Tcs[1], K);
    const double kappa = 4.0 * cst::onethird * cst::arad * cst::clight * ipow(p.Teff, 3) / rho / K.total;
    p.kappa = kappa;
    p.pres = P;
    dfdlnrho = 0.0;
    return P - p.tau * p.grav / kappa;
}

// bc09.f:607-666
double eval_pressure(double lnrho, double& dlnPdlnrho, void* ctx, int& ierr) {
    Par& p = *static_cast<Par*>(ctx);
    ierr = 0;
    const double rho = dsm::exp(lnrho);
    const double Tcs[3] = {0.0, 0.0, 0.0};
    double res[num_eos_results], chi = use_default_nuclear_size;
    int phase;
    eval_crust_eos(*p.h, p.eos, rho, p.temp, p.ionic, p.ncharged, p.Zion, p.Aion, p.Yion, Tcs, res, phase, chi);
    ++p.n_eos;
    if (!std::isfinite(res[i_lnP])) { ierr = -9; return 0.0; }
    dlnPdlnrho = res[i_chiRho];
    return res[i_lnP] - dsm::log(p.pres);
}

// bc09.f:553-605
double get_lnrho_from_PT(Par& p, double P, double T, double lnrho_guess, int& ierr) {
    p.temp = T;
    p.pres = P;
    double x1, x3, y1, y3;
    ierr = look_for_brackets(lnrho_guess, 0.5, x1, x3, eval_pressure, &p, y1, y3, p.st->max_iter_density);
    if (ierr != 0) return lnrho_guess;
    return safe_root_with_initial_guess(eval_pressure, &p, lnrho_guess, x1, x3, y1, y3, p.st->max_iter_density,
                                        p.st->tol_density, p.st->tol_density, ierr);
}

// bc09.f:485-551
int get_coefficients(Par& p, double P, double T, double& rho, double& kappa, double& del_ad) {
    int ierr = 0;
    const double lnrho = get_lnrho_from_PT(p, P, T, dsm::log(rho), ierr);
    if (ierr != 0) return ierr;
    rho = dsm::exp(lnrho);
    const double Tcs[3] = {0.0, 0.0, 0.0};
    double res[num_eos_results], chi = use_default_nuclear_size;
    int phase;
    eval_crust_eos(*p.h, p.eos, rho, T, p.ionic, p.ncharged, p.Zion, p.Aion, p.Yion, Tcs, res, phase, chi);
    ++p.n_eos;
    if (!(std::fabs(dsm::exp(res[i_lnP]) - P) < 1.0e-3 * P)) return -10;   // assertion "EOS is behaving" :536
    del_ad = res[i_grad_ad];
    CondComponents K;
    conductivity(p.cond, rho, T, chi, res[i_Gamma], res[i_Theta], res[i_mu_e], p.ionic, Tcs[1], K);
    kappa = 4.0 * cst::onethird * cst::arad * cst::clight * ipow(T, 3) / rho / K.total;
    return 0;
}

// bc09.f:445-483; returns ierr ("nonzero means retry with smaller timestep", :458)
int deriv(int, double lnP, const double* lnT4, double* dlnT4dlnP, void* ctx) {
    Par& p = *static_cast<Par*>(ctx);
    double rho = p.rho;
    const double P = dsm::exp(lnP);
    const double T = p.Teff * dsm::exp(0.25 * lnT4[0]);
    double kappa = 0.0, del_ad = 0.0;
    const int ierr = get_coefficients(p, P, T, rho, kappa, del_ad);
    if (ierr != 0) { dlnT4dlnP[0] = 0.0; return ierr; }
    double v = 0.75 * kappa * P / p.grav / dsm::exp(lnT4[0]);
    dlnT4dlnP[0] = std::min(v, 4.0 * del_ad);
    p.rho = rho;   // store the new anchor point
    return 0;
}

// bc09.f:284-371; rho_ph < 0 on input: ideal gas + Thomson guess
int find_photospheric_pressure(Par& p, double& rho_ph, double& P_ph, double& kappa) {
    const double Teff = p.Teff;
    const double fallback_Pphoto = 2.0 * cst::onethird * cst::arad * ipow(Teff, 4);
    const double kappa_Th = cst::Thomson * cst::avogadro * p.ionic.Ye;
    const double eps_ph = p.st->tol_photosphere_condition * p.tau * p.grav / kappa_Th;
    double lnrho_guess;
    if (rho_ph < 0.0) {
        const double Pgas = p.tau * p.grav / kappa_Th - cst::onethird * cst::arad * ipow(Teff, 4);
        if (Pgas < 0.0) { P_ph = fallback_Pphoto; kappa = kappa_Th; return -2; }
        lnrho_guess = dsm::log(Pgas * cst::amu * p.ionic.A / (p.ionic.Z + 1.0) / cst::boltzmann / Teff);
    } else {
        lnrho_guess = dsm::log(rho_ph);
    }
    double x1, x3, y1, y3;
    int ierr = look_for_brackets(lnrho_guess, 0.1, x1, x3, photosphere, &p, y1, y3, p.st->max_iter_photosphere);
    if (ierr != 0) {
        P_ph = fallback_Pphoto; rho_ph = dsm::exp(lnrho_guess); kappa = p.kappa;
        return ierr;
    }
    const double lnrho_ph = safe_root_with_initial_guess(photosphere, &p, lnrho_guess, x1, x3, y1, y3,
                                                         p.st->max_iter_photosphere, p.st->tol_photosphere_lnrho,
                                                         eps_ph, ierr);
    if (ierr != 0) {
        P_ph = fallback_Pphoto; rho_ph = dsm::exp(lnrho_guess); kappa = p.kappa;
        return ierr;
    }
    // (the reference takes P and kappa from the root finder's last function call, :367-369; here they are
    // evaluated at the returned root itself: the difference is below the root's own tolerance)
    double dummy;
    photosphere(lnrho_ph, dummy, &p, ierr);
    P_ph = p.pres;
    rho_ph = dsm::exp(lnrho_ph);
    kappa = p.kappa;
    return 0;
}

// bc09.f:97-282
int integrate_bc09_atm(Par& p, double lgyb, double lgy_light, double lgTeff, double& lgTb, double& rho,
                       Bc09Knot* info) {
    p.Teff = dexp10(lgTeff);
    p.tau = cst::twothird;
    set_layer(p, 0);
    double P, kappa;
    int ierr = find_photospheric_pressure(p, rho, P, kappa);
    if (info) { info->rho_ph = rho; info->P_ph = P; info->kappa_ph = kappa; }
    if (ierr != 0) return ierr;
    p.rho = rho; p.temp = p.Teff; p.pres = P; p.kappa = kappa;

    double lnP = dsm::log(P);
    double lnT4 = 0.0;
    const double lg_grav = dsm::log10(p.grav);
    double lnPend = std::max(cst::ln10 * (lgy_light + lg_grav), lnP + cst::ln10);
    lnPend = std::min(lnPend, cst::ln10 * (lgyb + lg_grav));
    Dop853Stats s1, s2;
    int idid = dop853_checked(1, deriv, &p, lnP, &lnT4, lnPend, 0.001, 0.0, 10000, 1.0e-6, 1.0e-6, &s1);
    if (idid < 0) return idid;
    set_layer(p, 1);
    lnPend = cst::ln10 * (lgyb + lg_grav);
    if (lnP < lnPend) {
        idid = dop853_checked(1, deriv, &p, lnP, &lnT4, lnPend, 0.001, 0.0, 10000, 1.0e-6, 1.0e-6, &s2);
        if (idid < 0) return idid;
    }
    lgTb = 0.25 * lnT4 / cst::ln10 + lgTeff;
    if (info) { info->nstep = s1.nstep + s2.nstep; info->nfcn = s1.nfcn + s2.nfcn; info->n_eos = (int)p.n_eos; }
    return 0;
}

}  // namespace

// bc09.f:45-95
int bc09_Teff(const HelmTable& h, double grav, double Plight, double Pb, int n, const double* lgTeff, double* lgTb,
              double* lgflux, const Bc09Settings& st, Bc09Knot* info) {
    Par p;
    p.h = &h;
    p.st = &st;
    p.cond.include_neutrons = false;
    p.cond.include_superfluid_phonons = false;
    p.cond.rad_full_off_lgrho = 10.0;
    p.cond.rad_full_on_lgrho = 9.0;
    p.grav = grav;
    const double lgyb = dsm::log10(Pb / grav);
    const double lgy_light = dsm::log10(Plight / grav);
    double rho_ph = -1.0;
    for (int i = n - 1; i >= 0; --i) {   // hottest first; the array is ascending
        if (!st.chain_guess) rho_ph = -1.0;
        p.n_eos = 0;
        const int ierr = integrate_bc09_atm(p, lgyb, lgy_light, lgTeff[i], lgTb[i], rho_ph, info ? info + i : nullptr);
        if (ierr != 0) return ierr;
        lgflux[i] = 4.0 * lgTeff[i] + dsm::log10(cst::sigma_SB);
    }
    return 0;
}

// dStar_atm_mod.f:46-106 with prefix 'bc09': knots lgTeff = 5.5 .. 7.0 (dStar_atm_def.f:7-8), table in lgTb
int bc09_table(const HelmTable& h, double grav, double Plight, double Pb, int nv, double* lgTb, double* lgTeff4,
               double* lgflux4, const Bc09Settings& st) {
    std::vector<double> lgTeff(nv), lgflux(nv);
    for (int i = 0; i < nv; ++i) lgTeff[i] = 5.5 + (double)i * (7.0 - 5.5) / (double)(nv - 1);
    const int ierr = bc09_Teff(h, grav, Plight, Pb, nv, lgTeff.data(), lgTb, lgflux.data(), st, nullptr);
    if (ierr != 0) return ierr;
    for (int i = 0; i < nv; ++i) { lgTeff4[4 * i] = lgTeff[i]; lgflux4[4 * i] = lgflux[i]; }
    interp_pm(lgTb, nv, lgTeff4);
    interp_pm(lgTb, nv, lgflux4);
    return 0;
}

}  // namespace dso
