// Device thermal conductivity (sm_100a, FP64).
//
// B200-native replacement for get_thermal_conductivity (conductivity/public/conductivity_lib.f:108-125)
// -> conductivity (conductivity/private/eval_conductivity.f:6-143): electron channels ee_SY06 /
// ee_PCY / eion / eQ / eQ_page (electron_conductivity.f:7-222), radiative opacity
// (photon_conductivity.f:7-108), neutron-impurity / neutron-phonon scattering
// (neutron_conductivity.f:69-383; two adaptive DOP853 quadratures per evaluation) and superfluid
// phonons sPh (:12-67).
//
// The impurity structure-factor integral does not depend on T (neutron_conductivity.f:382), so
// the coefficient-pass kernel evaluates it once per zone (ds_lambda_imp) and shares it across
// the 128 temperature lanes; the phonon integral is per lane.
#pragma once
#include "dconst.cuh"
#include "coef_tables.cuh"
#include "eos.cuh"

struct DsCondCtrl {  // conductivity_def.f:29-52
    int include_neutrons, include_sf_phonons, ee_fmla, eQ_fmla;
    double rad_full_off_lgrho, rad_full_on_lgrho;
};
struct DsCond {
    double total, ee, ei, eQ, sf, nQ, np, kap, electron_total, neutron_total;
};

namespace dsk {
using namespace dsc;


DS_D double ee_PCY(double ne, double T) {
    const double onesixth = (double)(1.0f / 6.0f);
    double yfac = ds_hc.yfac;
    double mec2 = Melectron * clight * clight;
    double eefac = 1.5 * (finestructure * finestructure) * mec2 / dsd::ipow<3>(pi) / hbar;
    double befac = ds_hc.befac;
    double kF = dsm::pow(3.0 * (pi * pi) * ne, onethird);
    double x = kF * hbar / Melectron / clight;
    double eF = sqrt(1.0 + x * x);
    double beta = x / eF;
    double b2 = beta * beta;
    double be3 = befac / dsm::pow(beta, 1.5);
    double plasma_theta = boltzmann * T / mec2;
    double y = yfac * dsm::pow(x, 1.5) / plasma_theta / sqrt(eF);
    double lJ = 1.0 + (DS_R4(2.810) - DS_R4(0.810) * b2) / y;
    double llJ = dsm::log(lJ);
    double J0 = onethird / dsd::ipow<3>(1.0 + DS_R4(0.07414) * y);
    double J1 = J0 * llJ + onesixth * dsd::ipow<5>(pi) * y / dsd::ipow<4>(DS_R4(13.91) + y);
    double J2 = 1.0 + DS_R4(0.4) * (3.0 + 1.0 / (x * x)) / (x * x);
    double J = dsd::ipow<3>(y) * J1 * J2;
    return eefac * (plasma_theta * plasma_theta) * J / eF / be3;
}

// ee_SY06 (electron_conductivity.f:44-92), split into the part that depends on the electron density only and the
// part that depends on T: the first holds a cube root, two square roots, two exponentials and ~10 divisions and is
// evaluated once per zone (DsCondPre); same operations in the same order as the one-piece form, so same bits.
struct DsEePre {
    double nupre_ne, eF, u, u2, u3, Tp, At, Ct2, Ct, c2404, Alt, Blt, Clt2, Clt, c1852, c182, inv_u;
};
DS_D DsEePre ee_SY06_pre(double ne) {
    DsEePre p;
    const double nu_pre = 36.0 * (finestructure * finestructure) * (hbar * hbar) * clight / pi / boltzmann / Melectron;
    const double Tp_pre = ds_hc.Tp_pre;
    const double kF = dsm::pow(threepisquare * ne, onethird);
    const double xF = hbar * kF / Melectron / clight;
    p.eF = sqrt(1.0 + xF * xF);
    p.u = xF / p.eF;
    p.Tp = Tp_pre * sqrt(ne / p.eF);
    p.u2 = p.u * p.u; p.u3 = p.u * p.u * p.u;
    const double u4 = dsd::ipow<4>(p.u);
    p.At = 20.0 + 450.0 * p.u3;
    const double Ct1 = DS_R4(0.05067) + DS_R4(0.03216) * p.u2;
    p.Ct2 = DS_R4(0.0254) + DS_R4(0.04127) * u4;
    p.Ct = p.At * dsm::exp(Ct1 / p.Ct2);
    p.c2404 = DS_R4(2.404) / p.Ct;
    p.Alt = DS_R4(12.2) + DS_R4(25.2) * p.u3;
    p.Blt = 1.0 - 0.75 * p.u;
    const double Clt1 = DS_R4(0.123636) + DS_R4(0.016234) * p.u2;
    p.Clt2 = DS_R4(0.0762) + DS_R4(0.05714) * u4;
    p.Clt = p.Alt * dsm::exp(Clt1 / p.Clt2);
    p.c1852 = DS_R4(18.52) * p.u2 / p.Clt;
    p.c182 = DS_R4(18.2) * p.u2 / p.Clt;
    p.inv_u = 1.0 / p.u;
    p.nupre_ne = nu_pre * ne;
    return p;
}
DS_D double ee_SY06_T(const DsEePre& p, double T) {
    const double theta = (double)sqrtf(3.0f) * p.Tp / T;
    const double tu = theta * p.u;
    double Il = p.inv_u * (DS_R4(0.1587) - DS_R4(0.02538) / (1.0 + DS_R4(0.0435) * theta)) *
                dsm::log(1.0 + DS_R4(128.56) / theta / (DS_R4(37.1) + theta * (DS_R4(10.83) + theta)));
    double It = p.u3 * (p.c2404 + (p.Ct2 - p.c2404) / (1.0 + DS_R4(0.1) * tu)) *
                dsm::log(1.0 + p.Ct / tu / (p.At + tu));
    double Ilt = p.u * (p.c1852 + (p.Clt2 - p.c182) / (1.0 + DS_R4(0.1558) * dsm::pow(theta, p.Blt))) *
                 dsm::log(1.0 + p.Clt / (p.Alt * theta + DS_R4(10.83) * (tu * tu) + dsm::pow(tu, (double)(8.0f / 3.0f))));
    double I = Il + It + Ilt;
    return p.nupre_ne * I / p.eF / T;
}
DS_D double ee_SY06(double ne, double T) { return ee_SY06_T(ee_SY06_pre(ne), T); }

DS_D double eone(double z) {
    const double a0 = DS_R4(0.1397), a1 = DS_R4(0.5772), a2 = DS_R4(2.2757);
    double z2 = z * z, z3 = z2 * z, z4 = z2 * z2;
    return dsm::exp(-z4 / (z3 + a0)) * (dsm::log(1.0 + 1.0 / z) - a1 / (1.0 + a2 * z2));
}

// eion (electron_conductivity.f:94-172), split the same way: what depends on (kF, composition) only -- a sixth
// root, a square root, a logarithm, ~10 divisions -- once per zone; what depends on Gamma_e and eta per temperature.
struct DsEiPre {
    double kF2, v2, s, eta02, aei2, beta, gsfac, gk01, fourkF2, w1, w_num, log1, sp1, s_sp1, log2, inv_sp1, r2s, z2a,
        eifac_eF;
};
DS_D DsEiPre eion_pre(double kF, double Z, double Z2, double A) {
    DsEiPre p;
    const double um2 = DS_R4(12.973);
    const double onesixth = (double)(1.0f / 6.0f), fourthird = 4.0 * onethird;
    const double electroncompton = hbar / Melectron / clight;
    const double eifac = fourthird * clight * (finestructure * finestructure) / electroncompton / pi;
    p.kF2 = kF * kF;
    const double x = electroncompton * kF;
    const double eF = sqrt(1.0 + x * x);
    const double v = x / eF;
    p.v2 = v * v;
    p.s = finestructure / pi / v;
    p.eta02 = dsd::ipow<2>(DS_R4(0.19) / dsm::pow(Z, onesixth));
    const double aei = ds_hc.aei_fac * kF;
    p.aei2 = aei * aei;
    p.beta = pi * finestructure * Z * v;
    p.gsfac = 1.0 + DS_R4(0.122) * (p.beta * p.beta);
    const double Gk1 = 1.0 + p.beta * dsd::ipow<3>(v);
    const double GkZ = 1.0 - 1.0 / Z;
    p.gk01 = DS_R4(0.0105) * GkZ * Gk1;
    p.fourkF2 = 4.0 * p.kF2;
    p.w1 = 1.0 + onethird * p.beta;
    p.w_num = 4.0 * um2 * p.kF2;
    p.log1 = dsm::log(1.0 + 1.0 / p.s);
    p.sp1 = p.s + 1.0;
    p.s_sp1 = p.s / (1.0 + p.s);
    p.log2 = dsm::log((p.s + 1.0) / p.s);
    p.inv_sp1 = 1.0 / (p.s + 1.0);
    p.r2s = (2.0 * p.s + 1.0) / (2.0 * p.s + 2.0);
    p.z2a = Z2 / A;
    p.eifac_eF = eifac * eF;
    return p;
}
DS_D double eion_T(const DsEiPre& p, double Gamma_e, double eta, double Z, double Z2, double Z53, double A) {
    const double um1 = DS_R4(2.8);
    const double sw_max = dbl_max_log;
    const double s = p.s;
    double Gamma = Gamma_e * Z53;
    (void)Gamma;
    double qD2 = 3.0 * Gamma_e * p.aei2 * Z2 / Z;
    double Gs = eta / sqrt(eta * eta + p.eta02) * p.gsfac;
    double Gk0 = 1.0 / dsm::pow(eta * eta + DS_R4(0.0081), 1.5);
    double Gk = Gs + p.gk01 * Gk0 * eta;
    double a0 = p.fourkF2 / qD2 / eta;
    double D1 = dsm::exp(DS_R4(-9.1) * eta);
    double D = dsm::exp(-a0 * um1 * D1 * 0.25);
    double w = p.w_num / qD2 * p.w1;
    double sw = s * w;
    double L1, L2;
    if (sw < sw_max) {
        double fac = dsm::exp(sw) * (eone(sw) - eone(sw + w));
        L1 = 0.5 * (p.log1 + (1.0 - dsm::exp(-w)) * s / p.sp1 - fac * (1.0 + sw));
        L2 = 0.5 * (1.0 - (1.0 - dsm::exp(-w)) / w +
                    s * (p.s_sp1 * (1.0 - dsm::exp(-w)) - 2.0 * p.log1 + fac * (2.0 + sw)));
    } else {
        L1 = 0.5 * (p.log2 - p.inv_sp1);
        L2 = p.r2s - s * p.log2;
    }
    double Lei = p.z2a * (L1 - p.v2 * L2) * Gk * D;
    return p.eifac_eF * Lei * A / Z;
}
DS_D double eion(double kF, double Gamma_e, double eta, double Ye, double Z, double Z2, double Z53, double A) {
    (void)Ye;
    return eion_T(eion_pre(kF, Z, Z2, A), Gamma_e, eta, Z, Z2, Z53, A);
}

DS_D double eQ(double kF, double Z, double A, double Q) {
    double electron_compton = hbar / Melectron / clight;
    double fac = 4.0 * onethird * clight * (finestructure * finestructure) / electron_compton / pi;
    double dZ2 = Q / A;
    double x = electron_compton * kF;
    double gamma = sqrt(1.0 + x * x);
    double v = x / gamma;
    double v2 = v * v;
    double s = finestructure / pi / v;
    double L1 = 1.0 + 4.0 * v2 * s;
    double L = 0.5 * (L1 * dsm::log(1.0 + 1.0 / s) - v2 - 1.0 / (s + 1.0) - v2 * s / (1.0 + s));
    return fac * gamma * dZ2 * L * A / Z;
}

DS_D double eQ_page(double kF, double Z, double A, double Q) {
    double electron_compton = hbar / Melectron / clight;
    double fac = 4.0 * onethird * clight * (finestructure * finestructure) / electron_compton / pi;
    double dZ2 = Q / A;
    double x = electron_compton * kF;
    double gamma = sqrt(1.0 + x * x);
    double v = x / gamma;
    double qs = DS_R4(0.048196) / sqrt(v);
    double L = dsm::log(1.0 / qs) - 0.5 * (1 + v * v);
    return fac * gamma * dZ2 * L * A / Z;
}

DS_D double electron_scattering(double eta_e, double theta, double Ye) {
    double xi = dsm::exp(DS_R4(0.8168) * eta_e - DS_R4(0.05522) * (eta_e * eta_e));
    double xi2 = xi * xi;
    double theta2 = theta * theta;
    double t1 = DS_R4(1.129) + DS_R4(0.2965) * xi - DS_R4(0.005594) * xi2;
    double t2 = DS_R4(11.47) + DS_R4(0.3570) * xi + DS_R4(0.1078) * xi2;
    double t3 = DS_R4(-3.249) + DS_R4(0.1678) * xi - DS_R4(0.04706) * xi2;
    double Gbar = t1 + t2 * theta + t3 * theta2;
    return Thomson * avogadro * Ye / Gbar;
}

DS_D double freefree(double rho, double T, double eta_e, const DsComp& c) {
    const double bfac = DS_R4(0.753);
    double rY = rho * c.Ye;
    double T8 = T * 1.0e-8;
    double T8_32 = dsm::pow(T8, 1.5);
    double xx;
    if (eta_e > dbl_max_log) xx = eta_e;
    else xx = dsm::log(1.0 + dsm::exp(eta_e));
    double norm = 1.16 * 8.02e3 * xx * T8_32 / rY;
    double x = dsm::pow(1.0 + xx, twothird);
    double gam = sqrt(DS_R4(1.58e-3) / T8) * c.Z;
    double sxpt = sqrt(x + 10.0);
    double sx = sqrt(x);
    double tpg = 2.0 * pi * gam;
    double expnum = dsm::exp(-tpg / sxpt);
    double expden = dsm::exp(-tpg / sx);
    double elwert = (1.0 - expnum) / (1.0 - expden);
    double relfactor = 1.0 + dsm::pow(T8 / DS_R4(7.7), 1.5);
    double gff = norm * elwert * relfactor;
    return bfac * (rY / 1.0e5) / dsm::pow(T8, 3.5) * gff * c.Z2 / c.A;
}

DS_D double Rosseland_kappa(double rho, double T, double mu_e, const DsComp& c) {
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

// ---------------------------------------------------------------- neutron channels
// Structure-factor integrands of Deibel et al. (2017), neutron_conductivity.f:291-383.
struct SfacPar {
    double T, Z, A, Yn, kfn, R_a;
};
template <bool PHONON>
DS_D double sfac(double x, const SfacPar& p) {
    double formfac = fabs(3.0 * (dsm::sin(x) - x * dsm::cos(x)) / dsd::ipow<3>(x));
    if (!PHONON) return dsd::ipow<3>(x) * (formfac * formfac);
    double nn = dsd::ipow<3>(p.kfn) / threepisquare;
    double nion = nn / p.Yn * (1.0 - p.Yn) / p.A;
    double aion = dsm::pow(3.0 / (4.0 * pi * nion), onethird);
    double mion = (p.A - p.Z) * Mneutron + p.Z * Mproton;
    double omegaP = dsm::pow(hbar * clight * 4. * pi * (p.Z * p.Z) * finestructure * nion / mion, 0.5);
    double eta = boltzmann * p.T / (hbar * omegaP);
    double Gamma = hbar * clight * (p.Z * p.Z) * finestructure / (boltzmann * p.T * aion);
    double a1 = dsd::ipow<2>(x / p.R_a * aion) / (3. * Gamma * eta);
    double um1 = DS_R4(2.8);
    double um2 = 13.0;
    double w = a1 * (um1 * dsm::exp(DS_R4(-9.1) * eta) + 2. * eta * um2) / 4.;
    double w1 = a1 * um2 * (eta * eta) / (2. * dsm::pow(eta * eta + dsd::ipow<2>(um2 / 117.), 0.5));
    double ssig = dsm::exp(-2. * (w - w1)) * (1.0 - dsm::exp(-2. * w1));
    double dskappa2 = a1 * 91. * (eta * eta) * dsm::exp(-2. * w) / dsd::ipow<2>(1. + DS_R4(111.4) * (eta * eta));
    double dskappa4 = a1 * DS_R4(0.101) * dsd::ipow<4>(eta) /
                      ((DS_R4(0.06408) + eta * eta) * dsm::pow(DS_R4(0.001377) + eta * eta, 1.5));
    double dskappa = dskappa2 + dskappa4;
    return dsd::ipow<3>(x) * (formfac * formfac) * (ssig + (3. * dsd::ipow<2>(p.kfn * p.R_a / x) - 0.5) * dskappa);
}

// Scalar DOP853 (Hairer, Norsett & Wanner; the integrator behind MESA's num_lib dop853 that
// neutron_conductivity.f:149,255 calls), specialised to one equation y' = g(x): rtol, atol scalar,
// h0 = 0 (automatic), no dense output.  Returns idid (1 ok, -2 too many steps, -3 step too small).
template <bool PHONON>
DS_D int dop853_quad(const SfacPar& par, double x, double xend, int max_steps, double rtol, double atol,
                     double& yout) {
    const double uround = 2.3e-16, safe = 0.9, facc1 = 1.0 / 0.333, facc2 = 1.0 / 6.0;
    double y = 0.0;
    const double hmax = fabs(xend - x);
    const double posneg = copysign(1.0, xend - x);
    double k[13];
    double facold = 1.0e-4;
    bool last = false, reject = false;
    int nstep = 0, naccpt = 0;
    k[0] = sfac<PHONON>(x, par);
    // hinit (iord = 8)
    double h;
    {
        double sk = atol + rtol * fabs(y);
        double dnf = (k[0] / sk) * (k[0] / sk), dny = (y / sk) * (y / sk);
        if (dnf <= 1.0e-10 || dny <= 1.0e-10) h = 1.0e-6; else h = sqrt(dny / dnf) * 0.01;
        h = fmin(h, hmax);
        h = copysign(h, posneg);
        double f1 = sfac<PHONON>(x + h, par);
        double der2 = ((f1 - k[0]) / sk) * ((f1 - k[0]) / sk);
        der2 = sqrt(der2) / h;
        double der12 = fmax(fabs(der2), sqrt(dnf));
        double h1;
        if (der12 <= 1.0e-15) h1 = fmax(1.0e-6, fabs(h) * 1.0e-3); else h1 = dsm::pow(0.01 / der12, 1.0 / 8);
        h = fmin(fmin(100.0 * fabs(h), h1), hmax);
        h = copysign(h, posneg);
    }
    int idid = 1;
    for (;;) {
        if (nstep > max_steps) { idid = -2; break; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { idid = -3; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
        ++nstep;
#pragma unroll
        for (int s = 1; s < 12; ++s) k[s] = sfac<PHONON>(x + dop_c[s] * h, par);  // y' = g(x): stages need no y
        double xph = x + h;
        double k4 = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) k4 += dop_b[j] * k[j];
        double ynew = y + h * k4;
        double sk = atol + rtol * fmax(fabs(y), fabs(ynew));
        double e3 = 0.0, e5 = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) { e3 += dop_e3[j] * k[j]; e5 += dop_e5[j] * k[j]; }
        double err2 = (e3 / sk) * (e3 / sk);
        double err = (e5 / sk) * (e5 / sk);
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (1 * deno));
        double fac11 = dsm::pow(err, 0.125);
        double fac = fac11 / dsm::pow(facold, 0.0);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0) {
            facold = fmax(err, 1.0e-4);
            ++naccpt;
            k[0] = sfac<PHONON>(xph, par);
            y = ynew;
            x = xph;
            if (last) break;
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
    yout = y;
    return idid;
}

struct DsNucleus {  // SLy4 nuclear size model, neutron_conductivity.f:130-135
    double n_in, R_a;
};
DS_D DsNucleus nucleus_model(const DsComp& c) {
    double isospin = 1.0 - 2.0 * c.Z / c.A;
    double n_0 = DS_R4(0.1740) / dsd::ipow<3>(fm_to_cm);
    double n_2 = DS_R4(-0.0157) / dsd::ipow<3>(fm_to_cm);
    double n_l = n_0 + n_2 * (isospin * isospin);
    DsNucleus r;
    r.n_in = n_l / 2.0 * (1.0 + DS_R4(0.9208) * isospin);
    r.R_a = dsm::pow(3.0 * (c.A - c.Z) / 4.0 / pi / r.n_in, onethird);
    return r;
}
// Lambda_n,imp: T-independent (neutron_conductivity.f:338-383, :149-153)
DS_D int ds_lambda_imp(double nn, const DsComp& c, double& lambda) {
    double kFn = dsm::pow(threepisquare * nn, onethird);
    DsNucleus nuc = nucleus_model(c);
    SfacPar par{0.0, c.Z, c.A, c.Yn, kFn, nuc.R_a};
    return dop853_quad<false>(par, 1.0e-10, 2.0 * kFn * nuc.R_a, 1000, DS_R4(1.0e-4), DS_R4(1.0e-5), lambda);
}
DS_D int ds_lambda_phonon(double nn, double T, const DsComp& c, double& lambda) {
    double kFn = dsm::pow(threepisquare * nn, onethird);
    DsNucleus nuc = nucleus_model(c);
    SfacPar par{T, c.Z, c.A, c.Yn, kFn, nuc.R_a};
    return dop853_quad<true>(par, 1.0e-10, 2.0 * kFn * nuc.R_a, 1000, DS_R4(1.0e-4), DS_R4(1.0e-5), lambda);
}
// common prefactor of n_imp (:155-164) and n_phonon (:261-269): nu = pre * [Q/Z^2] * Lambda
DS_D double neutron_nu_prefactor(double nn, double nion, const DsComp& c) {
    const double twothird4 = (double)(2.0f / 3.0f);
    DsNucleus nuc = nucleus_model(c);
    double fac = (double)(4.0f / 27.0f) / pi / dsd::ipow<3>(hbar) / (clight * clight);
    double V = hbar * hbar * dsm::pow(threepisquare * nuc.n_in, twothird4) / 2.0 / Mneutron *
               (1.0 - dsm::pow(nn / nuc.n_in, twothird4));
    double mnstar = Mneutron;
    double EFn = DS_R4(15.1) * mev_to_ergs * dsm::pow(c.Yn, twothird4) *
                     dsm::pow(nn * amu / c.Yn / DS_R4(1.0E14), twothird4) * (2.0 * Mneutron / mnstar) +
                 Mneutron * (clight * clight);
    return fac * EFn * (nion / nn) * dsd::ipow<2>(V * nuc.R_a);
}

// superfluid phonons, Aguilera et al. (2009); neutron_conductivity.f:12-67 (nn, nion in fm^-3)
DS_D double sPh(double nn, double nion, double temperature, const DsComp& c) {
    const double anl = DS_R4(10.0);
    double K_n = boltzmann * clight * fm_to_cm / dsd::ipow<3>(hbarc_n * fm_to_cm);
    double etwo = finestructure * hbarc_n;
    double Mu_n = amu * clight2 * ergs_to_mev;
    double T = temperature * boltzmann * ergs_to_mev;
    double ai = dsm::pow(3.0 / fourpi / nion, onethird);
    double kFn = dsm::pow(threepisquare * nn, onethird);
    double kFe = dsm::pow(threepisquare * nion * c.Z, onethird);
    double vs = hbarc_n * kFn / Mn_n / (double)sqrtf(3.0f);
    double Cv = 2.0 * (pi * pi) * dsd::ipow<3>(T) / 15.0 / dsd::ipow<3>(vs);
    double omega_p = sqrt(fourpi * etwo * c.Z2 * nion / c.A / Mu_n);
    double qTFe = ds_hc.yfac * kFe;
    double cs = omega_p / qTFe;
    double alpha = cs / vs;
    double anlt = anl / (1.0 + DS_R4(0.4) * kFn * anl + 0.5 * 2.0 * anl * (kFn * kFn));
    double gmix = 2.0 * anlt * hbarc_n * sqrt(nion * kFn / c.A / (Mu_n * Mu_n));
    double TUm = onethird * etwo * omega_p * dsm::pow(c.Z, onethird);
    double omega = 3.0 * T / hbarc_n;
    double fu = dsm::exp(-TUm / T);
    double Llphn = 2.0 / pi / omega;
    double Llphu = 100.0 * ai;
    double Llph = 1.0 / (fu / Llphu + (1.0 - fu) / Llphn);
    double tau_lph = Llph / cs;
    double wt = tau_lph * omega;
    double Lsph = Llph * dsd::ipow<2>(vs / gmix) * (1.0 + dsd::ipow<2>(1.0 - alpha * alpha) * (wt * wt)) / alpha / (wt * wt);
    return onethird * Cv * vs * Lsph * K_n;
}

// eval_conductivity.f:6-143.  lambda_imp: the zone's precomputed impurity integral, or NaN to have
// it evaluated here; imp_ierr its status.
// T-independent part of the conductivity of a zone (evaluated once per zone by ds_zone_aux_kernel<FACE>): log10 rho,
// the electron density / Fermi wavenumber / Fermi energy chain (a cube root and a square root) and the whole
// electron-impurity collision frequency, which depends on (kF, Z, A, Q) only (electron_conductivity.f:174-222).
struct DsCondPre {
    double lgrho, ne, kF, xF, eF, nu_eQ;
    DsEePre ee;   // ee_SY06, density part
    DsEiPre ei;   // eion, (kF, composition) part
};
DS_D DsCondPre ds_cond_pre(const DsCondCtrl& rq, double rho, const DsComp& c) {
    DsCondPre p;
    p.lgrho = dsm::log10(rho);
    p.ne = rho / amu * c.Ye;
    p.kF = dsm::pow(threepisquare * p.ne, onethird);
    p.xF = hbar * p.kF / Melectron / clight;
    p.eF = sqrt(1.0 + p.xF * p.xF);
    p.nu_eQ = (rq.eQ_fmla == 2) ? eQ_page(p.kF, c.Z, c.A, c.Q) : eQ(p.kF, c.Z, c.A, c.Q);
    p.ee = ee_SY06_pre(p.ne);
    p.ei = eion_pre(p.kF, c.Z, c.Z2, c.A);
    return p;
}

DS_D void ds_conductivity(const DsCondCtrl& rq, double rho, double T, double chi, double Gamma,
                          double eta, double mu_e, const DsComp& c, const DsCondPre& cp, double Tns, double lambda_imp,
                          int imp_ierr, DsCond& kap) {
    const double eQ_threshold = 1.0e-8, sf_reduction_threshold = 1.0e-10;
    kap = DsCond{0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double lgrho = cp.lgrho;
    const double ne = cp.ne;
    const double kF = cp.kF;
    const double eF = cp.eF;
    const double Gamma_e = Gamma / c.Z53;
    const double kappa_e_pre = onethird * (pi * pi) * (boltzmann * boltzmann) * T * ne / Melectron / eF;
    double nu = 0.0, nu_c;

    DS_PHASE();
    nu_c = (rq.ee_fmla == 2) ? ee_PCY(ne, T) : ee_SY06_T(cp.ee, T);
    nu = nu + nu_c;
    DS_PHASE();
    if (nu_c > dbl_tiny) kap.ee = kappa_e_pre / nu_c;
    nu_c = eion_T(cp.ei, Gamma_e, eta, c.Z, c.Z2, c.Z53, c.A);
    nu = nu + nu_c;
    if (nu_c > dbl_tiny) kap.ei = kappa_e_pre / nu_c;
    DS_PHASE();
    nu_c = cp.nu_eQ;   // eQ / eQ_page(kF, Z, A, Q): zone-level, see DsCondPre
    nu = nu + nu_c;
    if (c.Q > eQ_threshold) kap.eQ = kappa_e_pre / nu_c;
    kap.electron_total = kappa_e_pre / nu;
    DS_PHASE();

    if (c.Yn > 0.0) {
        const double nn = rho / amu * c.Yn / (1.0 - chi);
        const double nion = rho / amu * (1.0 - c.Yn) / c.A;
        if (rq.include_neutrons) {
            nu = 0.0;
            const double kappa_n_pre = onethird * (pi * pi) * (boltzmann * boltzmann) * T * nn / Mneutron;
            double R = 1.0;
            if (T < Tns) {
                double tau = T / Tns;
                double v = sqrt(1.0 - tau) * (DS_R4(1.456) - DS_R4(0.157) / sqrt(tau) + DS_R4(1.764) / tau);
                R = dsm::pow(DS_R4(0.4186) + sqrt((double)(1.007f * 1.007f) + dsd::ipow<2>(DS_R4(0.5010) * v)), 2.5) *
                    dsm::exp(DS_R4(1.456) - sqrt((double)(1.456f * 1.456f) + v * v));
                if (R < sf_reduction_threshold) R = 0.0;
            }
            const double pre = neutron_nu_prefactor(nn, nion, c);
            double lam = lambda_imp;
            int ierr = imp_ierr;
            if (isnan(lam)) ierr = ds_lambda_imp(nn, c, lam);
            if (ierr >= 0) {
                nu_c = pre * c.Q / (c.Z * c.Z) * lam;
                kap.nQ = kappa_n_pre / nu_c * R;
                nu = nu + nu_c;
            }
            ierr = ds_lambda_phonon(nn, T, c, lam);
            if (ierr >= 0) {
                nu_c = pre * lam;
                kap.np = kappa_n_pre / nu_c * R;
                nu = nu + nu_c;
            }
            if (nu > 0.0) kap.neutron_total = kappa_n_pre / nu * R;
        }
        if (rq.include_sf_phonons && T < Tns) kap.sf = sPh(nn / density_n, nion / density_n, T, c);
    }

    DS_PHASE();
    double K_opacity = 0.0;
    if (lgrho < rq.rad_full_off_lgrho) {
        kap.kap = Rosseland_kappa(rho, T, mu_e, c);
        if (lgrho > rq.rad_full_on_lgrho) {
            double alpha = (lgrho - rq.rad_full_on_lgrho) / (rq.rad_full_off_lgrho - rq.rad_full_on_lgrho);
            kap.kap = kap.kap * (1.0 - alpha);
        }
        K_opacity = 4.0 * onethird * arad * clight * dsd::ipow<3>(T) / rho / kap.kap;
    }
    kap.total = kap.electron_total + kap.neutron_total + kap.sf + K_opacity;
}
}  // namespace dsk
