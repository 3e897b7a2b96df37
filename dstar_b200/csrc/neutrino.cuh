// Device crust neutrino emissivities (sm_100a, FP64).
//
// B200-native replacement for get_crust_neutrino_emissivity (neutrino/public/neutrino_lib.f:16-30)
// -> emissivity (neutrino/private/eval_crust_neutrino.f:14-54): pair (:104-161), photo (:163-187)
// with the cfp fit (:189-245), plasma (:56-102), e-ion bremsstrahlung (:275-299) and 1S0
// pair-breaking (:247-273).  Result-compatible with the reference including two quirks:
// plasma() feeds T, not log10 T, into (x,y) so fxy == 1 in the crust; cfp() zeroes fp for
// c*xi > 100 and then overwrites it.
#pragma once
#include "dconst.cuh"
#include "coef_tables.cuh"

struct DsNeutrino {
    double total, pair, photo, plasma, brems, pbf;
};

namespace dsn {
using namespace dsc;
constexpr double two_thirds = dsc::twothird;
constexpr double cn_dn = 2.0;
constexpr double cn_snthtw = 0.2319;
constexpr double cn_cv = 0.5 + 2.0 * cn_snthtw;
constexpr double cn_ca = 0.5;
constexpr double cn_cvp = 1.0 - cn_cv;
constexpr double cn_cap = 1.0 - cn_ca;


template <bool FAST>
DS_D double plasma(double t, double r, const DsZonePre& zp, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double cvd = cn_cv * cn_cv + cn_dn * (cn_cvp * cn_cvp);
    double rdouble = zp.pl_lg;   // log10(2 r): zone-level, see DsZonePre
    double x = (17.5 + rdouble - 3.0 * t) / den_c(6.0);
    double y = (-24.5 + rdouble + 3.0 * t) / den_c(6.0);
    double enum_ = fmin(0.0, y - 1.6 + 1.25 * x);
    double fxy;
    if (fabs(x) > 0.7 || y < 0.0) {
        fxy = 1.0;
    } else {
        fxy = 1.05 + (0.39 - 1.25 * x - 0.35 * dsm::sin(4.5 * x) - 0.3 * dsm::exp(-dsd::ipow<2>(4.5 * x + 0.9))) *
                         dsm::exp(-dsd::ipow<2>(enum_ / den(0.57 - 0.25 * x)));
    }
    double lamb = 1.686e-10 * t;
    double gamd = 1.1095e11 * r / den((t * t) * sq(1.0 + zp.pl_p23));
    double gam = sq(gamd);
    double ft = 2.4 + 0.6 * sq(gam) + 0.51 * gam + 1.25 * dsm::pow(gam, 1.5);
    double fl = (8.6 * gamd + 1.35 * dsm::pow(gam, 7.0 / den_c(2.0))) / den(225.0 - 17.0 * gam + gamd);
    double qap;
    if (fabs(sq(gamd)) > 700.0) qap = 0.0;
    else qap = fxy * (fl + ft) * dsm::exp(-gam) * dsd::ipow<3>(gamd) * dsd::ipow<9>(lamb) * 3.0e21;
    return qap * cvd;
}

template <bool FAST>
DS_D double pair(double temp, double rhom, const DsZonePre& zp, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double a0 = 6.002e19, a1 = 2.084e20, a2 = 1.872e21;
    const double b1 = 9.383e-1, b2 = -4.141e-1, b3 = 5.829e-2, c = 5.5924e0;
    const double bh1 = 1.2383, bh2 = -0.8141, bh3 = 0.0, ch = 4.9924;
    const double coep = (cn_cv * cn_cv + cn_ca * cn_ca) + cn_dn * (cn_cvp * cn_cvp + cn_cap * cn_cap);
    const double coem = (cn_cv * cn_cv - cn_ca * cn_ca) + cn_dn * (cn_cvp * cn_cvp - cn_cap * cn_cap);
    double lam = temp / den_c(5.9302e9);
    double xi = zp.cbrt_rY9 / den(lam);
    double qp1 = 10.74800 * (lam * lam) + 0.3967 * sq(lam) + 1.005;
    double qp2 = rhom / den(7.692e7 * dsd::ipow<3>(lam) + 9.715e6 * sq(lam));
    double qp = (1.0 / den(qp1)) * dsm::pow(1.0 + qp2, -0.3);
    double axi = a0 + a1 * xi + a2 * (xi * xi);
    double fp1, fp2;
    if (xi > 100.0) fp1 = 0.0;
    else if (temp < 1.0e10) fp1 = axi * dsm::exp(-c * xi);
    else fp1 = axi * dsm::exp(-ch * xi);
    if (temp < 1.0e10) fp2 = dsd::ipow<3>(xi) + b1 / den(lam) + b2 / den(lam * lam) + b3 / den(dsd::ipow<3>(lam));
    else fp2 = dsd::ipow<3>(xi) + bh1 / den(lam) + bh2 / den(lam * lam) + bh3 / den(dsd::ipow<3>(lam));
    double fp = fp1 / den(fp2);
    double g = 1.0 - 13.04 * (lam * lam) + 133.5 * dsd::ipow<4>(lam) + 1534.0 * dsd::ipow<6>(lam) + 918.6 * dsd::ipow<8>(lam);
    if (lam < 7.0e-3) return 0.0;
    return 0.5 * coep * (1.0 + coem / den(coep) * qp) * g * dsm::exp(-2.0 / den(lam)) * fp;
}

template <bool FAST>
DS_D double cfp(double lamb, double xi, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double b[3] = {6.290e-3, 7.483e-3, 3.061e-4};
    double temp = lamb * 5.9302e9;
    int band;
    double c, tau;
    if (temp < 1.0e7) return 0.0;
    else if (temp < 1.0e8) { band = 0; c = 0.5654 + dsm::log10(temp / den_c(1.0e7)); tau = dsm::log10(temp / den_c(1.0e7)); }
    else if (temp < 1.0e9) { band = 1; c = 1.5654; tau = dsm::log10(temp / den_c(1.0e8)); }
    else { band = 2; c = 1.5654; tau = dsm::log10(temp / den_c(1.0e9)); }
    // the harmonics do not depend on j: evaluate the six arguments once (one fused sincos each) and
    // reuse them for a(0), a(1), a(2); accumulation order as eval_crust_neutrino.f:230-237
    const double c10 = dsm::cos(10.0 * pi * tau);
    double ck[6], sk[6];
#pragma unroll
    for (int k = 1; k <= 5; ++k) {
        double pjt = pi * (double)k * tau;
        dsm::sincos(5.0 / den_c(3.0) * pjt, sk[k], ck[k]);
    }
    double a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        a[j] = 0.5 * cfp_cc[band][j][0] + 0.5 * cfp_cc[band][j][6] * c10;
#pragma unroll
        for (int k = 1; k <= 5; ++k) a[j] = a[j] + cfp_cc[band][j][k] * ck[k] + cfp_dd[band][j][k - 1] * sk[k];
    }
    return (a[0] + a[1] * xi + a[2] * (xi * xi)) * dsm::exp(-c * xi) /
           (dsd::ipow<3>(xi) + b[0] / den(lamb) + b[1] / den(lamb * lamb) + b[2] / den(dsd::ipow<3>(lamb)));
}

template <bool FAST>
DS_D double photo(double rhom, double lamb, const DsZonePre& zp, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double coe1 = 0.5 * ((cn_cv * cn_cv + cn_ca * cn_ca) + cn_dn * (cn_cvp * cn_cvp + cn_cap * cn_cap));
    const double coe2 = ((cn_cv * cn_cv - cn_ca * cn_ca) + cn_dn * (cn_cvp * cn_cvp - cn_cap * cn_cap)) /
                        ((cn_cv * cn_cv + cn_ca * cn_ca) + cn_dn * (cn_cvp * cn_cvp + cn_cap * cn_cap));
    double poly = 1.875e8 * lamb + 1.653e8 * (lamb * lamb) + 8.499e8 * dsd::ipow<3>(lamb) - 1.604e8 * dsd::ipow<4>(lamb);
    double qp = 0.666 * dsm::pow(1.0 + 2.045 * lamb, -2.066) * (1.0 / den(1.0 + rhom * (1.0 / poly)));
    double xi = zp.cbrt_rY9 / den(lamb);
    double fp = cfp<FAST>(lamb, xi, bad);
    return coe1 * (1.0 - coe2 * qp) * rhom * dsd::ipow<5>(lamb) * fp;
}

template <bool FAST>
DS_D double pbf(double kF, double T, double Tns, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double qpbf_pre = 1.17e21;
    double tau = T / Tns;  // Tns = 0 without a gap: plain division (T/0 = inf takes the early return)
    if (kF == 0.0 || tau > 1.0 || tau < 0.01) return 0.0;
    double xF = kF * hbarc_n / den_c(Mn_n);
    double eF = sq(1.0 + xF * xF);
    double mstar = eF;
    double v = sq(1.0 - tau) * (1.456 - 0.157 / den(sq(tau)) + 1.764 / den(tau));
    double v2 = v * v;
    double F1 = v2 * (0.602 + v2 * (0.5942 + 0.288 * v2));
    double F2 = sq(DS_R4(0.5547) + sq(0.4453 * 0.4453 + 0.01130 * v2));
    double F3 = dsm::exp(-sq(4.0 * v2 + 2.245 * 2.245) + 2.245);
    double F = F1 * F2 * F3;
    double reduct = dsd::ipow<2>(xF / den(eF));
    return qpbf_pre * (cn_dn + 1) * mstar * xF * dsd::ipow<7>(T / den_c(1.0e9)) * F * reduct;
}

template <bool FAST>
DS_D double bremsstrahlung(double rho, double T, const DsComp& c, const DsZonePre& zp, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double threefourth = 3.0 / 4.0, GF = 1.436e-49, Cp2 = 1.675;
    double qbrems_pre = 8.0 * pi * (GF * GF) * (finestructure * finestructure) * Cp2 *
                        dsd::ipow<6>(boltzmann / den_c(hbar) / den_c(clight)) / den_c(567.0) / den_c(hbar) / den_c(amu);
    double kF = zp.kFb;   // (3 pi^2 rho Ye / amu)^(1/3); re and eta likewise zone-level, see DsZonePre
    double tau = 0.5 * boltzmann * T / den_c(hbar) / den_c(clight) / den(kF);
    double eta = zp.etab;
    double Zbareta = c.Z * eta;
    double A = 0.269 + 20.0 * tau + 0.0168 * c.Z + 0.00121 * eta - 0.0356 * Zbareta + 0.0137 * c.Z2 * tau +
               1.54 * Zbareta * tau;
    double B = 1.0 + 180.0 * (tau * tau) + 0.483 * tau * c.Z + 20.0 * tau * Zbareta * eta + 4.31e-5 * c.Z2;
    double L = A / den(dsm::pow(B, threefourth));
    return qbrems_pre * dsd::ipow<6>(T) * c.Z2 * rho * L * (1.0 - c.Yn) / den(c.A);
}

template <bool FAST>
DS_D void ds_crust_neutrino(double rho, double T, const DsComp& c, const DsZonePre& zp, double chi, double Tcn,
                               const int* ch, DsNeutrino& eps, bool& bad) {
    DS_DEN_SCOPE(FAST);
    double Tme = Melectron * clight2 / den_c(boltzmann);
    double rY = rho * c.Ye;
    double lambda = T / den(Tme);
    eps = DsNeutrino{0, 0, 0, 0, 0, 0};
    DS_PHASE();
    if (ch[0]) eps.pair = pair<FAST>(T, rY, zp, bad);
    DS_PHASE();
    if (ch[1]) eps.photo = photo<FAST>(rY, lambda, zp, bad);
    DS_PHASE();
    if (ch[2]) eps.plasma = plasma<FAST>(T, rY, zp, bad);
    DS_PHASE();
    if (ch[3]) eps.brems = bremsstrahlung<FAST>(rho, T, c, zp, bad);
    DS_PHASE();
    if (ch[4]) {
        (void)chi;
        eps.pbf = pbf<FAST>(zp.kn_pbf, T, Tcn, bad);   // kn = (1.5 pi^2 nn)^(1/3) / cm_to_fm: zone-level
    }
    DS_PHASE();
    eps.total = eps.pair + eps.photo + eps.plasma + eps.brems + eps.pbf;
}

}  // namespace dsn
