// ORACLE (test infrastructure): thermal conductivity and the DOP853 quadrature it needs.
// Restates conductivity/private/eval_conductivity.f:6-143, electron_conductivity.f:7-222,
// neutron_conductivity.f:12-383, photon_conductivity.f:7-108.  DOP853 follows Hairer,
// Norsett & Wanner's dop853.f (the algorithm MESA's num_lib wraps; MESA source not in tree).
#include <algorithm>
#include <cmath>
#include <limits>

#include "dstar_oracle.hpp"

namespace dso {
using namespace cst;

// ---------------------------------------------------------------- DOP853
#include "dop853_coef.inc"

namespace {
template <class F>
double hinit853(int n, F f, double x, const double* y, double posneg,
                const double* f0, double* f1, double* y1, int iord, double hmax, double atol,
                double rtol) {
    double dnf = 0.0, dny = 0.0;
    for (int i = 0; i < n; ++i) {
        double sk = atol + rtol * std::fabs(y[i]);
        dnf += (f0[i] / sk) * (f0[i] / sk);
        dny += (y[i] / sk) * (y[i] / sk);
    }
    double h;
    if (dnf <= 1.0e-10 || dny <= 1.0e-10) h = 1.0e-6;
    else h = std::sqrt(dny / dnf) * 0.01;
    h = std::min(h, hmax);
    h = std::copysign(h, posneg);
    for (int i = 0; i < n; ++i) y1[i] = y[i] + h * f0[i];
    f(n, x + h, y1, f1);
    double der2 = 0.0;
    for (int i = 0; i < n; ++i) {
        double sk = atol + rtol * std::fabs(y[i]);
        der2 += ((f1[i] - f0[i]) / sk) * ((f1[i] - f0[i]) / sk);
    }
    der2 = std::sqrt(der2) / h;
    double der12 = std::max(std::fabs(der2), std::sqrt(dnf));
    double h1;
    if (der12 <= 1.0e-15) h1 = std::max(1.0e-6, std::fabs(h) * 1.0e-3);
    else h1 = dsm::pow(0.01 / der12, 1.0 / iord);
    h = std::min(std::min(100.0 * std::fabs(h), h1), hmax);
    return std::copysign(h, posneg);
}
}  // namespace

// returns idid: 1 ok, -2 too many steps, -3 step too small, -5 the right-hand side failed at the starting point.
// F(n, x, y, dy) returns ierr; MESA's fcn interface says "nonzero means retry with smaller timestep" (quoted at
// dStar_atm/private/bc09.f:458): a stage whose right-hand side fails is treated as a rejected step with the largest
// step reduction the controller knows (h / facc1), like a step with infinite error.
namespace {
template <class F>
int dop853_impl(int n, F f, double& x, double* y, double xend, double h0,
                double max_step, int max_steps, double rtol, double atol, Dop853Stats* st) {
    const double uround = 2.3e-16, safe = 0.9, fac1 = 0.333, fac2 = 6.0, beta = 0.0;
    const double expo1 = 1.0 / 8.0 - beta * 0.2;
    const double facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    double hmax = (max_step == 0.0) ? (xend - x) : max_step;
    hmax = std::fabs(hmax);
    double posneg = std::copysign(1.0, xend - x);
    std::vector<double> K(13 * (size_t)n), ynew(n), ytmp(n);
    auto k = [&](int s) { return K.data() + (size_t)s * n; };
    double facold = 1.0e-4;
    bool last = false, reject = false;
    int nfcn = 0, nstep = 0, naccpt = 0, nrejct = 0;
    if (f(n, x, y, k(0)) != 0) return -5;
    double h = h0;
    if (h == 0.0) h = hinit853(n, f, x, y, posneg, k(0), k(1), k(2), 8, hmax, atol, rtol);
    nfcn += 2;
    int idid = 1;
    for (;;) {
        if (nstep > max_steps) { idid = -2; break; }
        if (0.1 * std::fabs(h) <= std::fabs(x) * uround) { idid = -3; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
        ++nstep;
        // 12 stages
        bool failed = false;
        for (int s = 1; s < 12 && !failed; ++s) {
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                for (int j = 0; j < s; ++j) {
                    double a = dop_a[s][j];
                    if (a != 0.0) acc += a * k(j)[i];
                }
                ytmp[i] = y[i] + h * acc;
            }
            failed = f(n, x + dop_c[s] * h, ytmp.data(), k(s)) != 0;
        }
        nfcn += 11;
        if (failed) {   // retry with a smaller step
            h = h / facc1;
            reject = true;
            if (naccpt >= 1) ++nrejct;
            last = false;
            continue;
        }
        double xph = x + h;
        double* k4 = k(12);
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int j = 0; j < 12; ++j)
                if (dop_b[j] != 0.0) acc += dop_b[j] * k(j)[i];
            k4[i] = acc;
            ynew[i] = y[i] + h * acc;
        }
        double err = 0.0, err2 = 0.0;
        for (int i = 0; i < n; ++i) {
            double sk = atol + rtol * std::max(std::fabs(y[i]), std::fabs(ynew[i]));
            double e3 = 0.0, e5 = 0.0;
            for (int j = 0; j < 12; ++j) {
                if (dop_e3[j] != 0.0) e3 += dop_e3[j] * k(j)[i];
                if (dop_e5[j] != 0.0) e5 += dop_e5[j] * k(j)[i];
            }
            err2 += (e3 / sk) * (e3 / sk);
            err += (e5 / sk) * (e5 / sk);
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = std::fabs(h) * err * std::sqrt(1.0 / (n * deno));
        double fac11 = dsm::pow(err, expo1);
        double fac = fac11 / dsm::pow(facold, beta);
        fac = std::max(facc2, std::min(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0 && f(n, xph, ynew.data(), k4) != 0) {   // the new point itself is not evaluable: retry smaller
            ++nfcn;
            h = h / facc1;
            reject = true;
            if (naccpt >= 1) ++nrejct;
            last = false;
            continue;
        }
        if (err <= 1.0) {
            facold = std::max(err, 1.0e-4);
            ++naccpt;
            ++nfcn;
            for (int i = 0; i < n; ++i) { k(0)[i] = k4[i]; y[i] = ynew[i]; }
            x = xph;
            if (last) { h = hnew; idid = 1; break; }
            if (std::fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * std::min(std::fabs(hnew), std::fabs(h));
            reject = false;
        } else {
            hnew = h / std::min(facc1, fac11 / safe);
            reject = true;
            if (naccpt >= 1) ++nrejct;
            last = false;
        }
        h = hnew;
    }
    if (st) { st->nfcn = nfcn; st->nstep = nstep; st->naccpt = naccpt; st->nrejct = nrejct; }
    return idid;
}
}  // namespace
int dop853(int n, OdeRhs f, void* ctx, double& x, double* y, double xend, double h0,
           double max_step, int max_steps, double rtol, double atol, Dop853Stats* st) {
    auto g = [f, ctx](int n_, double x_, const double* y_, double* dy_) { f(n_, x_, y_, dy_, ctx); return 0; };
    return dop853_impl(n, g, x, y, xend, h0, max_step, max_steps, rtol, atol, st);
}
int dop853_checked(int n, OdeRhsChecked f, void* ctx, double& x, double* y, double xend, double h0,
                   double max_step, int max_steps, double rtol, double atol, Dop853Stats* st) {
    auto g = [f, ctx](int n_, double x_, const double* y_, double* dy_) { return f(n_, x_, y_, dy_, ctx); };
    return dop853_impl(n, g, x, y, xend, h0, max_step, max_steps, rtol, atol, st);
}

// ---------------------------------------------------- electron scattering
namespace {
// electron_conductivity.f:7-42
double ee_PCY(double ne, double T) {
    const double onesixth = (double)(1.0f / 6.0f);
    double yfac = std::sqrt(4.0 * finestructure / pi);
    double mec2 = Melectron * clight * clight;
    double eefac = 1.5 * (finestructure * finestructure) * mec2 / ipow(pi, 3) / hbar;
    double befac = dsm::pow(finestructure / pi, 1.5);
    double kF = dsm::pow(3.0 * (pi * pi) * ne, onethird);
    double x = kF * hbar / Melectron / clight;
    double eF = std::sqrt(1.0 + x * x);
    double beta = x / eF;
    double b2 = beta * beta;
    double be3 = befac / dsm::pow(beta, 1.5);
    double plasma_theta = boltzmann * T / mec2;
    double y = yfac * dsm::pow(x, 1.5) / plasma_theta / std::sqrt(eF);
    double lJ = 1.0 + (R4(2.810) - R4(0.810) * b2) / y;
    double llJ = dsm::log(lJ);
    double J0 = onethird / ipow(1.0 + R4(0.07414) * y, 3);
    double J1 = J0 * llJ + onesixth * ipow(pi, 5) * y / ipow(R4(13.91) + y, 4);
    double J2 = 1.0 + R4(0.4) * (3.0 + 1.0 / (x * x)) / (x * x);
    double J = ipow(y, 3) * J1 * J2;
    return eefac * (plasma_theta * plasma_theta) * J / eF / be3;
}

// electron_conductivity.f:44-92
double ee_SY06(double ne, double T) {
    double nu_pre = 36.0 * (finestructure * finestructure) * (hbar * hbar) * clight / pi / boltzmann / Melectron;
    double Tp_pre = hbar * std::sqrt(4.0 * pi * finestructure * hbar * clight / Melectron) / boltzmann;
    double kF = dsm::pow(threepisquare * ne, onethird);
    double xF = hbar * kF / Melectron / clight;
    double eF = std::sqrt(1.0 + xF * xF);
    double u = xF / eF;
    double Tp = Tp_pre * std::sqrt(ne / eF);
    double theta = (double)std::sqrt(3.0f) * Tp / T;
    double u2 = u * u, u3 = u * u * u, u4 = ipow(u, 4);
    double tu = theta * u;
    double At = 20.0 + 450.0 * u3;
    double Ct1 = R4(0.05067) + R4(0.03216) * u2;
    double Ct2 = R4(0.0254) + R4(0.04127) * u4;
    double Ct = At * dsm::exp(Ct1 / Ct2);
    double Alt = R4(12.2) + R4(25.2) * u3;
    double Blt = 1.0 - 0.75 * u;
    double Clt1 = R4(0.123636) + R4(0.016234) * u2;
    double Clt2 = R4(0.0762) + R4(0.05714) * u4;
    double Clt = Alt * dsm::exp(Clt1 / Clt2);
    double Il = (1.0 / u) * (R4(0.1587) - R4(0.02538) / (1.0 + R4(0.0435) * theta)) *
                dsm::log(1.0 + R4(128.56) / theta / (R4(37.1) + theta * (R4(10.83) + theta)));
    double It = u3 * (R4(2.404) / Ct + (Ct2 - R4(2.404) / Ct) / (1.0 + R4(0.1) * tu)) *
                dsm::log(1.0 + Ct / tu / (At + tu));
    double Ilt = u * (R4(18.52) * u2 / Clt + (Clt2 - R4(18.2) * u2 / Clt) / (1.0 + R4(0.1558) * dsm::pow(theta, Blt))) *
                 dsm::log(1.0 + Clt / (Alt * theta + R4(10.83) * (tu * tu) + dsm::pow(tu, (double)(8.0f / 3.0f))));
    double I = Il + It + Ilt;
    return nu_pre * ne * I / eF / T;
}

// electron_conductivity.f:163-171 (internal function of eion)
double eone(double z) {
    const double a0 = R4(0.1397), a1 = R4(0.5772), a2 = R4(2.2757);
    double z2 = z * z, z3 = z2 * z, z4 = z2 * z2;
    return dsm::exp(-z4 / (z3 + a0)) * (dsm::log(1.0 + 1.0 / z) - a1 / (1.0 + a2 * z2));
}

// electron_conductivity.f:94-161
double eion(double kF, double Gamma_e, double eta, double Ye, double Z, double Z2, double Z53, double A) {
    (void)Ye;
    const double um1 = R4(2.8), um2 = R4(12.973);
    const double onesixth = (double)(1.0f / 6.0f), fourthird = 4.0 * onethird;
    double sw_max = dsm::log(std::numeric_limits<double>::max());
    double electroncompton = hbar / Melectron / clight;
    double eifac = fourthird * clight * (finestructure * finestructure) / electroncompton / pi;
    double kF2 = kF * kF;
    double x = electroncompton * kF;
    double eF = std::sqrt(1.0 + x * x);
    double v = x / eF;
    double v2 = v * v;
    double Gamma = Gamma_e * Z53;
    double s = finestructure / pi / v;
    double kTF2 = 4.0 * s * kF2;
    double eta02 = ipow(R4(0.19) / dsm::pow(Z, onesixth), 2);
    double aei = dsm::pow((double)(4.0f / 9.0f) / pi, onethird) * kF;
    double qD2 = 3.0 * Gamma_e * (aei * aei) * Z2 / Z;
    double beta = pi * finestructure * Z * v;
    double qi2 = qD2 * (1.0 + R4(0.06) * Gamma) * dsm::exp(-std::sqrt(Gamma));
    double qs2 = (qi2 + kTF2) * dsm::exp(-beta);
    (void)qs2;
    double Gs = eta / std::sqrt(eta * eta + eta02) * (1.0 + R4(0.122) * (beta * beta));
    double Gk0 = 1.0 / dsm::pow(eta * eta + R4(0.0081), 1.5);
    double Gk1 = 1.0 + beta * ipow(v, 3);
    double GkZ = 1.0 - 1.0 / Z;
    double Gk = Gs + R4(0.0105) * GkZ * Gk1 * Gk0 * eta;
    double a0 = 4.0 * kF2 / qD2 / eta;
    double D1 = dsm::exp(R4(-9.1) * eta);
    double D = dsm::exp(-a0 * um1 * D1 * 0.25);
    double w1 = 1.0 + onethird * beta;
    double w = 4.0 * um2 * kF2 / qD2 * w1;
    double sw = s * w;
    double L1, L2;
    if (sw < sw_max) {
        double fac = dsm::exp(sw) * (eone(sw) - eone(sw + w));
        L1 = 0.5 * (dsm::log(1.0 + 1.0 / s) + (1.0 - dsm::exp(-w)) * s / (s + 1.0) - fac * (1.0 + sw));
        L2 = 0.5 * (1.0 - (1.0 - dsm::exp(-w)) / w +
                    s * (s / (1.0 + s) * (1.0 - dsm::exp(-w)) - 2.0 * dsm::log(1.0 + 1.0 / s) + fac * (2.0 + sw)));
    } else {
        L1 = 0.5 * (dsm::log((s + 1.0) / s) - 1.0 / (s + 1.0));
        L2 = (2.0 * s + 1.0) / (2.0 * s + 2.0) - s * dsm::log((s + 1.0) / s);
    }
    double Lei = (Z2 / A) * (L1 - v2 * L2) * Gk * D;
    return eifac * eF * Lei * A / Z;
}

// electron_conductivity.f:174-200
double eQ(double kF, double Z, double A, double Q) {
    double electron_compton = hbar / Melectron / clight;
    double fac = 4.0 * onethird * clight * (finestructure * finestructure) / electron_compton / pi;
    double dZ2 = Q / A;
    double x = electron_compton * kF;
    double gamma = std::sqrt(1.0 + x * x);
    double v = x / gamma;
    double v2 = v * v;
    double s = finestructure / pi / v;
    double L1 = 1.0 + 4.0 * v2 * s;
    double L = 0.5 * (L1 * dsm::log(1.0 + 1.0 / s) - v2 - 1.0 / (s + 1.0) - v2 * s / (1.0 + s));
    return fac * gamma * dZ2 * L * A / Z;
}

// electron_conductivity.f:202-222
double eQ_page(double kF, double Z, double A, double Q) {
    double electron_compton = hbar / Melectron / clight;
    double fac = 4.0 * onethird * clight * (finestructure * finestructure) / electron_compton / pi;
    double dZ2 = Q / A;
    double x = electron_compton * kF;
    double gamma = std::sqrt(1.0 + x * x);
    double v = x / gamma;
    double qs = R4(0.048196) / std::sqrt(v);
    double L = dsm::log(1.0 / qs) - 0.5 * (1 + v * v);
    return fac * gamma * dZ2 * L * A / Z;
}

// ------------------------------------------------------- photon opacity
// photon_conductivity.f:7-26
double electron_scattering(double eta_e, double theta, double Ye) {
    double xi = dsm::exp(R4(0.8168) * eta_e - R4(0.05522) * (eta_e * eta_e));
    double xi2 = xi * xi;
    double theta2 = theta * theta;
    double t1 = R4(1.129) + R4(0.2965) * xi - R4(0.005594) * xi2;
    double t2 = R4(11.47) + R4(0.3570) * xi + R4(0.1078) * xi2;
    double t3 = R4(-3.249) + R4(0.1678) * xi - R4(0.04706) * xi2;
    double Gbar = t1 + t2 * theta + t3 * theta2;
    return Thomson * avogadro * Ye / Gbar;
}

// photon_conductivity.f:32-78
double freefree(double rho, double T, double eta_e, const Composition& c) {
    const double bfac = R4(0.753);
    double rY = rho * c.Ye;
    double T8 = T * 1.0e-8;
    double T8_32 = dsm::pow(T8, 1.5);
    double xx;
    if (eta_e > dsm::log(std::numeric_limits<double>::max())) xx = eta_e;
    else xx = dsm::log(1.0 + dsm::exp(eta_e));
    double norm = 1.16 * 8.02e3 * xx * T8_32 / rY;
    double x = dsm::pow(1.0 + xx, twothird);
    double gam = std::sqrt(R4(1.58e-3) / T8) * c.Z;
    double sxpt = std::sqrt(x + 10.0);
    double sx = std::sqrt(x);
    double tpg = 2.0 * pi * gam;
    double expnum = dsm::exp(-tpg / sxpt);
    double expden = dsm::exp(-tpg / sx);
    double elwert = (1.0 - expnum) / (1.0 - expden);
    double relfactor = 1.0 + dsm::pow(T8 / R4(7.7), 1.5);
    double gff = norm * elwert * relfactor;
    return bfac * (rY / 1.0e5) / dsm::pow(T8, 3.5) * gff * c.Z2 / c.A;
}

// photon_conductivity.f:80-108
double Rosseland_kappa(double rho, double T, double mu_e, const Composition& c) {
    double eta_e = mu_e * mev_to_ergs / boltzmann / T;
    double theta = T * boltzmann / Melectron / clight2;
    double kap_th = electron_scattering(eta_e, theta, c.Ye);
    double kap_ff = freefree(rho, T, eta_e, c);
    double T6 = T * 1.0e-6;
    double T_Ry = T6 / 0.15789 / c.Z2;
    double f = kap_ff / (kap_ff + kap_th);
    double A = 1.0 + (1.097 + 0.777 * T_Ry) / (1.0 + 0.536 * T_Ry) * dsm::pow(f, 0.617) * dsm::pow(1.0 - f, 0.77);
    return (kap_th + kap_ff) * A;
}

// ------------------------------------------------------ neutron channels
struct SfacPar { double T, Z, A, Yn, kfn, R_a; };

// common part of neutron_conductivity.f:291-336 / :338-383
inline void sfac_common(double x, const SfacPar& p, double& ssig, double& dskappa, double& formfac) {
    double nn = ipow(p.kfn, 3) / threepisquare;
    double nion = nn / p.Yn * (1.0 - p.Yn) / p.A;
    double aion = dsm::pow(3.0 / (4.0 * pi * nion), onethird);
    double mion = (p.A - p.Z) * Mneutron + p.Z * Mproton;
    double omegaP = dsm::pow(hbar * clight * 4. * pi * (p.Z * p.Z) * finestructure * nion / mion, (double)(1.f / 2.f));
    double eta = boltzmann * p.T / (hbar * omegaP);
    double Gamma = hbar * clight * (p.Z * p.Z) * finestructure / (boltzmann * p.T * aion);
    double a1 = ipow(x / p.R_a * aion, 2) / (3. * Gamma * eta);
    double um1 = R4(2.8);
    double um2 = 13.0;
    double w = a1 * (um1 * dsm::exp(R4(-9.1) * eta) + 2. * eta * um2) / 4.;
    double w1 = a1 * um2 * (eta * eta) /
                (2. * dsm::pow(eta * eta + ipow(um2 / 117., 2), (double)(1.f / 2.f)));
    ssig = dsm::exp(-2. * (w - w1)) * (1.0 - dsm::exp(-2. * w1));
    double dskappa2 = a1 * 91. * (eta * eta) * dsm::exp(-2. * w) / ipow(1. + R4(111.4) * (eta * eta), 2);
    double dskappa4 = a1 * R4(0.101) * ipow(eta, 4) /
                      ((R4(0.06408) + eta * eta) * dsm::pow(R4(0.001377) + eta * eta, 1.5));
    dskappa = dskappa2 + dskappa4;
    formfac = std::fabs(3.0 * (dsm::sin(x) - x * dsm::cos(x)) / ipow(x, 3));
}
void sfac_phonon(int, double x, const double*, double* dy, void* ctx) {
    const SfacPar& p = *static_cast<const SfacPar*>(ctx);
    double ssig, dskappa, formfac;
    sfac_common(x, p, ssig, dskappa, formfac);
    dy[0] = ipow(x, 3) * (formfac * formfac) * (ssig + (3. * ipow(p.kfn * p.R_a / x, 2) - 0.5) * dskappa);
}
void sfac_impurity(int, double x, const double*, double* dy, void* ctx) {
    const SfacPar& p = *static_cast<const SfacPar*>(ctx);
    double ssig, dskappa, formfac;
    sfac_common(x, p, ssig, dskappa, formfac);
    dy[0] = ipow(x, 3) * (formfac * formfac);
}

// neutron_conductivity.f:282-289
double neutron_potential(double nn, double n_in) {
    const double twothird4 = (double)(2.0f / 3.0f);
    return hbar * hbar * dsm::pow(threepisquare * n_in, twothird4) / 2.0 / Mneutron *
           (1.0 - dsm::pow(nn / n_in, twothird4));
}

// shared front of n_imp (:69-174) and n_phonon (:176-280)
double neutron_scattering(double nn, double nion, double temperature, const Composition& c,
                          bool impurity, int& ierr) {
    const double twothird4 = (double)(2.0f / 3.0f);
    ierr = 0;
    double kFn = dsm::pow(threepisquare * nn, onethird);
    double rtol = R4(1.0e-4), atol = R4(1.0e-5);
    double isospin = 1.0 - 2.0 * c.Z / c.A;
    double n_0 = R4(0.1740) / ipow(fm_to_cm, 3);
    double n_2 = R4(-0.0157) / ipow(fm_to_cm, 3);
    double n_l = n_0 + n_2 * (isospin * isospin);
    double n_in = n_l / 2.0 * (1.0 + R4(0.9208) * isospin);
    double R_a = dsm::pow(3.0 * (c.A - c.Z) / 4.0 / pi / n_in, onethird);
    double y = 0.0;
    double kn_start = 1.0e-10;
    double kn_end = 2.0 * kFn * R_a;
    SfacPar par{temperature, c.Z, c.A, c.Yn, kFn, R_a};
    double xx = kn_start;
    int idid = dop853(1, impurity ? sfac_impurity : sfac_phonon, &par, xx, &y, kn_end, 0.0, 0.0, 1000,
                      rtol, atol);
    if (idid < 0) ierr = idid;
    double Lambda = y;
    double fac = (double)(4.0f / 27.0f) / pi / ipow(hbar, 3) / (clight * clight);
    double V = neutron_potential(nn, n_in);
    double mnstar = Mneutron;
    double EFn = R4(15.1) * mev_to_ergs * dsm::pow(c.Yn, twothird4) *
                     dsm::pow(nn * amu / c.Yn / R4(1.0E14), twothird4) * (2.0 * Mneutron / mnstar) +
                 Mneutron * (clight * clight);
    if (impurity) return fac * EFn * (nion / nn) * ipow(V * R_a, 2) * c.Q / (c.Z * c.Z) * Lambda;
    return fac * EFn * (nion / nn) * ipow(V * R_a, 2) * Lambda;
}

// neutron_conductivity.f:12-67
double sPh(double nn, double nion, double temperature, const Composition& c) {
    const double anl = R4(10.0);
    double K_n = boltzmann * clight * fm_to_cm / ipow(hbarc_n * fm_to_cm, 3);
    double etwo = finestructure * hbarc_n;
    double Mu_n = amu * clight2 * ergs_to_mev;
    double T = temperature * boltzmann * ergs_to_mev;
    double ai = dsm::pow(3.0 / fourpi / nion, onethird);
    double kFn = dsm::pow(threepisquare * nn, onethird);
    double kFe = dsm::pow(threepisquare * nion * c.Z, onethird);
    double vs = hbarc_n * kFn / Mn_n / (double)std::sqrt(3.0f);
    double Cv = 2.0 * (pi * pi) * ipow(T, 3) / 15.0 / ipow(vs, 3);
    double omega_p = std::sqrt(fourpi * etwo * c.Z2 * nion / c.A / Mu_n);
    double qTFe = std::sqrt(4.0 * finestructure / pi) * kFe;
    double cs = omega_p / qTFe;
    double alpha = cs / vs;
    double anlt = anl / (1.0 + R4(0.4) * kFn * anl + 0.5 * 2.0 * anl * (kFn * kFn));
    double gmix = 2.0 * anlt * hbarc_n * std::sqrt(nion * kFn / c.A / (Mu_n * Mu_n));
    double TUm = onethird * etwo * omega_p * dsm::pow(c.Z, onethird);
    double omega = 3.0 * T / hbarc_n;
    double fu = dsm::exp(-TUm / T);
    double Llphn = 2.0 / pi / omega;
    double Llphu = 100.0 * ai;
    double Llph = 1.0 / (fu / Llphu + (1.0 - fu) / Llphn);
    double tau_lph = Llph / cs;
    double wt = tau_lph * omega;
    double Lsph = Llph * ipow(vs / gmix, 2) * (1.0 + ipow(1.0 - alpha * alpha, 2) * (wt * wt)) / alpha / (wt * wt);
    return onethird * Cv * vs * Lsph * K_n;
}
}  // namespace

// eval_conductivity.f:6-143
void conductivity(const CondControls& rq, double rho, double T, double chi, double Gamma,
                  double eta, double mu_e, const Composition& c, double Tns, CondComponents& kap) {
    const double eQ_threshold = 1.0e-8, sf_reduction_threshold = 1.0e-10;
    const double tiny = std::numeric_limits<double>::min();
    kap = CondComponents{0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double lgrho = dsm::log10(rho);
    double ne = rho / amu * c.Ye;
    double kF = dsm::pow(threepisquare * ne, onethird);
    double xF = hbar * kF / Melectron / clight;
    double eF = std::sqrt(1.0 + xF * xF);
    double Gamma_e = Gamma / c.Z53;
    double kappa_e_pre = onethird * (pi * pi) * (boltzmann * boltzmann) * T * ne / Melectron / eF;
    double nu = 0.0, nu_c;

    nu_c = (rq.ee_scattering_fmla == 2) ? ee_PCY(ne, T) : ee_SY06(ne, T);
    nu = nu + nu_c;
    if (nu_c > tiny) kap.ee = kappa_e_pre / nu_c;

    nu_c = eion(kF, Gamma_e, eta, c.Ye, c.Z, c.Z2, c.Z53, c.A);
    nu = nu + nu_c;
    if (nu_c > tiny) kap.ei = kappa_e_pre / nu_c;

    nu_c = (rq.eQ_scattering_fmla == 2) ? eQ_page(kF, c.Z, c.A, c.Q) : eQ(kF, c.Z, c.A, c.Q);
    nu = nu + nu_c;
    if (c.Q > eQ_threshold) kap.eQ = kappa_e_pre / nu_c;

    kap.electron_total = kappa_e_pre / nu;

    if (c.Yn > 0.0) {
        double nn = rho / amu * c.Yn / (1.0 - chi);
        double nion = rho / amu * (1.0 - c.Yn) / c.A;
        if (rq.include_neutrons) {
            nu = 0.0;
            double mnstar = Mneutron;
            double kappa_n_pre = onethird * (pi * pi) * (boltzmann * boltzmann) * T * nn / mnstar;
            double R = 1.0;
            if (T < Tns) {
                double tau = T / Tns;
                double v = std::sqrt(1.0 - tau) * (R4(1.456) - R4(0.157) / std::sqrt(tau) + R4(1.764) / tau);
                R = dsm::pow(R4(0.4186) + std::sqrt((double)(1.007f * 1.007f) + ipow(R4(0.5010) * v, 2)), 2.5) *
                    dsm::exp(R4(1.456) - std::sqrt((double)(1.456f * 1.456f) + v * v));
                if (R < sf_reduction_threshold) R = 0.0;
            }
            int ierr;
            nu_c = neutron_scattering(nn, nion, T, c, true, ierr);
            if (ierr != 0) nu_c = 0.0;
            else { kap.nQ = kappa_n_pre / nu_c * R; nu = nu + nu_c; }
            nu_c = neutron_scattering(nn, nion, T, c, false, ierr);
            if (ierr != 0) nu_c = 0.0;
            else { kap.np = kappa_n_pre / nu_c * R; nu = nu + nu_c; }
            if (nu > 0.0) kap.neutron_total = kappa_n_pre / nu * R;
        }
        if (rq.include_superfluid_phonons && T < Tns)
            kap.sf = sPh(nn / density_n, nion / density_n, T, c);
    }

    double K_opacity = 0.0;
    kap.kap = 0.0;
    if (lgrho < rq.rad_full_off_lgrho) {
        kap.kap = Rosseland_kappa(rho, T, mu_e, c);
        if (lgrho > rq.rad_full_on_lgrho) {
            double alpha = (lgrho - rq.rad_full_on_lgrho) / (rq.rad_full_off_lgrho - rq.rad_full_on_lgrho);
            kap.kap = kap.kap * (1.0 - alpha);
        }
        K_opacity = 4.0 * onethird * arad * clight * ipow(T, 3) / rho / kap.kap;
    }
    kap.total = kap.electron_total + kap.neutron_total + kap.sf + K_opacity;
}

}  // namespace dso
