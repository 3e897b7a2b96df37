// Device crust equation of state (sm_100a, FP64).
//
// B200-native replacement for eval_crust_eos (dStar_eos/public/dStar_eos_lib.f:123-297) and the
// components it sums: lattice_properties (:84-98), nuclear_volume_fraction (:100-109),
// neutron_wavenumber (:111-121), ee_exchange (private/electron.f:53-250), ion_mixture /
// LM_corrections / one_ion / ie_liquid / ie_lattice / ii_liquid / ii_lattice (private/ion.f:20-576),
// MB77 (private/neutron.f:15-69) with Antia's Fermi-integral fits (private/funct_fermi1.f),
// get_radiation_eos (private/radiation.f:5-25).  The HELM electron table lookup is in helm.cuh.
//
// Everything is inlined into the calling kernel so that outputs a kernel does not consume (free
// energy, internal energy, entropy in the coefficient pass) are dead-code-eliminated by nvcc.
// Constants whose value comes out of a libm call at start-up in the reference (cube roots of pi
// etc.) are evaluated once on the host and held in constant memory (DsHostConsts) so that both
// sides of the C ABI see identical bits.
#pragma once
#include "dconst.cuh"
#include "helm.cuh"
#include "coef_tables.cuh"

struct DsThermo {
    double f, u, p, s, cv, dpr, dpt;
};
struct DsEosCtrl {  // dStar_eos_private_def.f:12-21
    double Gamma_melt, rsi_melt, Ythresh;
};
// host-evaluated numeric constants (see ds_fill_host_consts in capi.cu)
struct DsHostConsts {
    double theta_fac;   // 2.0*(4.0/9.0/pi)**(2.0/3.0)         electron.f:56
    double ctf_fac;     // (18.0/175.0)*(12.0/pi)**(2.0*onethird) ion.f:184
    double xrs;         // finestructure*(9.0*pi/4.0)**onethird  ion.f:198,303
    double atf;         // (54.0/175.0)*(12.0/pi)**onethird*finestructure ion.f:301
    double e1;          // dsm::exp(1.0_dp) ion.f:300
    double half_log_6_over_pi;  // 0.5*dsm::log(6.0/pi) ion.f:432
    double aei_fac;     // (4.0/9.0/pi)**onethird electron_conductivity.f:128
    double befac;       // (finestructure/pi)**1.5 electron_conductivity.f:27
    double yfac;        // sqrt(4.0*finestructure/pi) electron_conductivity.f:24
    double Tp_pre;      // hbar*sqrt(4 pi alpha hbar c/me)/kB electron_conductivity.f:60
};
__constant__ DsHostConsts ds_hc;

constexpr int ds_liquid_phase = 1, ds_solid_phase = 2;
#define DS_DEFAULT_NUCLEAR_RADIUS DS_R4(1.13)

namespace dse {
using namespace dsc;


template <bool FAST>
DS_D void lattice_properties(double rho, double T, const DsComp& c, const DsZonePre& zp, double& ne, double& rs,
                        double& Gamma_e, double& Gamma, double& ionQ, bool& bad) {
    DS_DEN_SCOPE(FAST);
    (void)rho;
    ne = zp.ne;   // rho Ye / amu and (3 / 4 pi ne)^(1/3) / a_Bohr: zone-level, see DsZonePre
    rs = zp.rs;
    Gamma_e = 2.0 * Rydberg / den_c(boltzmann) / den(T) / den(rs);
    Gamma = Gamma_e * c.Z53;
    ionQ = sq(3.0 * (Gamma_e * Gamma_e) * Melectron / den(rs) / den_c(amu) * c.Z2XoA2 * c.A / den(c.Z));
}

DS_D double nuclear_volume_fraction(double rho, const DsComp& c, double nuclear_radius) {
    return onethird * fourpi * dsd::ipow<3>(nuclear_radius * fm_to_cm) * (rho * (1.0 - c.Yn) / amu);
}

DS_D double neutron_wavenumber(double rho, const DsComp& c, double chi) {
    double n = rho * c.Yn / amu / (1.0 - chi);
    return dsm::pow(threepisquare * n, onethird) / cm_to_fm;
}

// Functions of the nuclear charge alone that the ion-electron fits need (11 full pow/log evaluations per species
// and call).  They do not depend on the zone or on T, so the host evaluates them ONCE per species of the network
// with the same dsmath and the same expression order -- bit-identical by construction (tests/test_dsmath.py) --
// and the kernels read them by broadcast (all lanes of a group work on the same species at the same time).
struct DsIonZ {
    double Z53, Z76;                                     // one_ion   (ion.f:152-153)
    double Z13, cDH, cTF, lZ, aL, bL, nuL, q1L, q2L;     // ie_liquid (ion.f:178-285)
    double r3, b0, b1, b2, b3, Z23;                      // ie_lattice (ion.f:287-405)
};
DS_HD DsIonZ ds_ionz_compute(double Z, double ctf_fac) {
    DsIonZ q;
    const double sevensixth = (double)(7.0f / 6.0f);
    q.Z53 = dsm::pow(Z, dsc::fivethird);
    q.Z76 = dsm::pow(Z, sevensixth);
    const double s3 = (double)sqrtf(3.0f);
    q.Z13 = dsm::pow(Z, dsc::onethird);
    q.cDH = (Z / s3) * (dsm::pow(1.0 + Z, 1.5) - 1.0 - dsm::pow(Z, 1.5));
    q.cTF = ctf_fac * (Z * Z) * (q.Z13 - 1.0 + DS_R4(0.2) / sqrt(q.Z13));
    q.lZ = dsm::log(Z);
    q.aL = DS_R4(1.11) * dsm::pow(Z, DS_R4(0.475));
    q.bL = DS_R4(0.2) + DS_R4(0.078) * (q.lZ * q.lZ);
    q.nuL = DS_R4(1.16) + DS_R4(0.08) * q.lZ;
    q.q1L = DS_R4(0.18) / dsm::pow(Z, 0.25);
    q.q2L = DS_R4(0.2) + DS_R4(0.37) / sqrt(Z);
    const double a[4] = {DS_R4(1.1866), DS_R4(0.684), DS_R4(17.9), DS_R4(41.5)};
    q.r3 = 1.0 / (1.0 + DS_R4(0.01) * dsm::pow(q.lZ, 1.5) + DS_R4(0.097) / (Z * Z));
    q.b0 = 1.0 - a[0] / dsm::pow(Z, DS_R4(0.267)) + DS_R4(0.27) / Z;
    q.b1 = 1.0 + 2.25 / dsm::pow(Z, dsc::onethird) * (1.0 + a[1] * dsd::ipow<5>(Z) + DS_R4(0.222) * dsd::ipow<6>(Z)) /
                     (1.0 + DS_R4(0.222) * dsd::ipow<6>(Z));
    q.b2 = a[3] / (1.0 + q.lZ);
    q.b3 = DS_R4(0.395) * q.lZ + DS_R4(0.347) / dsm::pow(Z, 1.5);
    q.Z23 = dsm::pow(Z, dsc::twothird);
    return q;
}

template <bool FAST>
DS_D void ee_exchange(double Gamma_e, double rs, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    // theta_fac = 2.0*(4.0/9.0/pi)**(2.0/3.0): 4.0/9.0 and 2.0/3.0 are REAL(4) operations
    const double theta_fac = ds_hc.theta_fac;
    const double theta_low = DS_R4(0.005);
    const DsDen<FAST> Gamma_eD = den(Gamma_e);
    double theta = theta_fac * rs / Gamma_eD;
    double sqth = sq(theta);
    double theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta3 * theta;
    const DsDen<FAST> thetaD = den(theta), theta2D = den(theta2);
    double t1, t2, t1_h, t2_h, t1_hh, t2_hh;
    if (theta > theta_low) {
        double cht1 = dsm::cosh(1.0 / thetaD), sht1 = dsm::sinh(1.0 / thetaD);
        double cht2 = dsm::cosh(1.0 / den(sqth)), sht2 = dsm::sinh(1.0 / den(sqth));
        t1 = sht1 / den(cht1);
        t2 = sht2 / den(cht2);
        // (theta cosh(1/theta))^3 reaches 1e260 just above theta_low: these four stay plain divisions
        t1_h = -1.0 / dsd::ipow<2>(theta * cht1);
        t1_hh = 2.0 / dsd::ipow<3>(theta * cht1) * (cht1 - sht1 / thetaD);
        t2_h = -0.5 * sqth / dsd::ipow<2>(theta * cht2);
        t2_hh = (0.75 * sqth * cht2 - 0.5 * sht2) / dsd::ipow<3>(theta * cht2);
    } else {
        t1 = 1.0; t2 = 1.0; t1_h = 0.0; t2_h = 0.0; t1_hh = 0.0; t2_hh = 0.0;
    }
    const DsDen<FAST> t1D = den(t1), t2D = den(t2);
    const DsDen<FAST> a0 = den(0.75 + DS_R4(3.04363) * theta2 - DS_R4(0.09227) * theta3 + DS_R4(1.7035) * theta4);
    double a0_h = DS_R4(6.08726) * theta - DS_R4(0.27681) * theta2 + DS_R4(6.814) * theta3;
    double a0_hh = DS_R4(6.08726) - DS_R4(0.55362) * theta + DS_R4(20.442) * theta2;
    const DsDen<FAST> a1 = den(1. + DS_R4(8.31051) * theta2 + DS_R4(5.1105) * theta4);
    double a1_h = DS_R4(16.62102) * theta + DS_R4(20.442) * theta3;
    double a1_hh = DS_R4(16.62102) + DS_R4(61.326) * theta2;
    double a = DS_R4(0.610887) * a0 / a1 * t1;
    double ah = a0_h / a0 - a1_h / a1 + t1_h / t1D;
    double a_h = a * ah;
    double a_hh = a_h * ah + a * (a0_hh / a0 - dsd::ipow<2>(a0_h / a0) - a1_hh / a1 + dsd::ipow<2>(a1_h / a1) +
                                  t1_hh / t1D - dsd::ipow<2>(t1_h / t1D));

    DS_PHASE_INNER();
    const DsDen<FAST> b0 = den(DS_R4(0.341308) + DS_R4(12.070873) * theta2 + DS_R4(1.148889) * theta4);
    double b0_h = DS_R4(24.141746) * theta + DS_R4(4.595556) * theta3;
    double b0_hh = DS_R4(24.141746) + DS_R4(13.786668) * theta2;
    const DsDen<FAST> b1 = den(1.0 + DS_R4(10.495346) * theta2 + DS_R4(1.326623) * theta4);
    double b1_h = DS_R4(20.990692) * theta + DS_R4(5.306492) * theta3;
    double b1_hh = DS_R4(20.990692) + DS_R4(15.919476) * theta2;
    double b = sqth * t2 * b0 / b1;
    double bh = 0.5 / thetaD + t2_h / t2D + b0_h / b0 - b1_h / b1;
    double b_h = b * bh;
    double b_hh = b_h * bh + b * (-0.5 / theta2D + t2_hh / t2D - dsd::ipow<2>(t2_h / t2D) + b0_hh / b0 -
                                  dsd::ipow<2>(b0_h / b0) - b1_hh / b1 + dsd::ipow<2>(b1_h / b1));

    const DsDen<FAST> d0 = den(DS_R4(0.614925) + DS_R4(16.996055) * theta2 + DS_R4(1.489056) * theta4);
    double d0_h = DS_R4(33.99211) * theta + DS_R4(5.956224) * theta3;
    double d0_hh = DS_R4(33.99211) + DS_R4(17.868672) * theta2;
    const DsDen<FAST> d1 = den(1.0 + DS_R4(10.10935) * theta2 + DS_R4(1.22184) * theta4);
    double d1_h = DS_R4(20.2187) * theta + DS_R4(4.88736) * theta3;
    double d1_hh = DS_R4(20.2187) + DS_R4(14.66208) * theta2;
    double d = sqth * t2 * d0 / d1;
    double dh = 0.5 / thetaD + t2_h / t2D + d0_h / d0 - d1_h / d1;
    double d_h = d * dh;
    double d_hh = d_h * dh + d * (-0.5 / theta2D + t2_hh / t2D - dsd::ipow<2>(t2_h / t2D) + d0_hh / d0 -
                                  dsd::ipow<2>(d0_h / d0) - d1_hh / d1 + dsd::ipow<2>(d1_h / d1));

    DS_PHASE_INNER();
    const DsDen<FAST> e0 = den(DS_R4(0.539409) + DS_R4(2.522206) * theta2 + DS_R4(0.178484) * theta4);
    double e0_h = DS_R4(5.044412) * theta + DS_R4(0.713936) * theta3;
    double e0_hh = DS_R4(5.044412) + DS_R4(2.141808) * theta2;
    const DsDen<FAST> e1 = den(1.0 + DS_R4(2.555501) * theta2 + DS_R4(0.146319) * theta4);
    double e1_h = DS_R4(5.111002) * theta + DS_R4(0.585276) * theta3;
    double e1_hh = DS_R4(5.111002) + DS_R4(1.755828) * theta2;
    const DsDen<FAST> e = den(theta * t1 * e0 / e1);
    const DsDen<FAST> e2D = den(e * e);
    double eh = 1.0 / thetaD + t1_h / t1D + e0_h / e0 - e1_h / e1;
    double e_h = e * eh;
    double e_hh = e_h * eh + e * (t1_hh / t1D - dsd::ipow<2>(t1_h / t1D) + e0_hh / e0 - dsd::ipow<2>(e0_h / e0) -
                                  e1_hh / e1 + dsd::ipow<2>(e1_h / e1) - 1.0 / theta2D);

    double exp1th = dsm::exp(-1.0 / thetaD);
    const DsDen<FAST> c = den((DS_R4(0.872496) + DS_R4(0.025248) * exp1th) * e);
    // exp(-1/theta) runs down to 1e-300 in the crust: its two quotients stay plain divisions (a branch-free one
    // would flag every lane with theta in (0.0013, 0.0017) and send the whole CTA through the re-evaluation)
    double c_h = DS_R4(0.025248) * exp1th / theta2 * e + c * e_h / e;
    double c_hh = DS_R4(0.025248) * exp1th / theta2 * (e_h + (1.0 - 2.0 * theta) / theta2D * e) +
                  c_h * e_h / e + c * e_hh / e - c * dsd::ipow<2>(e_h / e);
    const DsDen<FAST> discr = den(sq(4.0 * e - d * d));
    double di_h = 0.5 / discr * (4.0 * e_h - 2.0 * d * d_h);
    double di_hh = (-dsd::ipow<2>((2.0 * e_h - d * d_h) / discr) + 2.0 * e_hh - d_h * d_h - d * d_hh) / discr;

    DS_PHASE_INNER();
    double s1 = -c / e * Gamma_e;
    double s1h = c_h / c - e_h / e;
    double s1_h = s1 * s1h;
    double s1_hh = s1_h * s1h + s1 * (c_hh / c - dsd::ipow<2>(c_h / c) - e_hh / e + dsd::ipow<2>(e_h / e));
    double s1_g = -c / e;
    double s1_hg = s1_g * (c_h / c - e_h / e);
    DS_PHASE_INNER();
    const DsDen<FAST> b2 = den(b - c * d / e);
    double b2_h = b_h - (c_h * d + c * d_h) / e + c * d * e_h / e2D;
    double b2_hh = b_hh - (c_hh * d + 2.0 * c_h * d_h + c * d_hh) / e +
                   (2.0 * (c_h * d + c * d_h - c * d * e_h / e) * e_h + c * d * e_hh) / e2D;

    const DsDen<FAST> sqge = den(sq(Gamma_e));
    double s2 = -2.0 / e * b2 * sqge;
    double s2h = b2_h / b2 - e_h / e;
    double s2_h = s2 * s2h;
    double s2_hh = s2_h * s2h + s2 * (b2_hh / b2 - dsd::ipow<2>(b2_h / b2) - e_hh / e + dsd::ipow<2>(e_h / e));
    double s2_g = 0.5 * s2 / Gamma_eD;
    double s2_gg = -0.5 * s2_g / Gamma_eD;
    double s2_hg = 0.5 * s2_h / Gamma_eD;

    const DsDen<FAST> r3 = den(e * Gamma_e + d * sqge + 1.0);
    double r3_h = e_h * Gamma_e + d_h * sqge;
    double r3_hh = e_hh * Gamma_e + d_hh * sqge;
    double r3_g = e + 0.5 * d / sqge;
    double r3_gg = -0.25 * d / den(Gamma_e * sqge);
    double r3_hg = e_h + 0.5 * d_h / sqge;

    DS_PHASE_INNER();
    double b3 = a - c / e;
    double b3_h = a_h - c_h / e + c * e_h / e2D;
    double b3_hh = a_hh - c_hh / e + (2.0 * c_h * e_h + c * e_hh) / e2D - 2.0 * c * (e_h * e_h) / den(dsd::ipow<3>(e));

    const DsDen<FAST> c3 = den((d / e * b2 - b3) / e);
    double c3_h = (d_h * b2 + d * b2_h + b3 * e_h) / e2D - 2.0 * d * b2 * e_h / den(dsd::ipow<3>(e)) - b3_h / e;
    double c3_hh = (-b3_hh + (d_hh * b2 + 2.0 * d_h * b2_h + d * b2_hh + b3_h * e_h + b3 * e_hh + b3_h * e_h) / e -
                    2.0 * ((d_h * b2 + d * b2_h + b3 * e_h + d_h * b2 + d * b2_h) * e_h + d * b2 * e_hh) / e2D +
                    6.0 * d * b2 * (e_h * e_h) / den(dsd::ipow<3>(e))) / e;

    double s3 = c3 * dsm::log(r3);
    double s3_h = s3 * c3_h / c3 + c3 * r3_h / r3;
    double s3_hh = (s3_h * c3_h + s3 * c3_hh) / c3 - s3 * dsd::ipow<2>(c3_h / c3) +
                   (c3_h * r3_h + c3 * r3_hh) / r3 - c3 * dsd::ipow<2>(r3_h / r3);
    double s3_g = c3 * r3_g / r3;
    double s3_gg = c3 * (r3_gg / r3 - dsd::ipow<2>(r3_g / r3));
    double s3_hg = (c3_h * r3_g + c3 * r3_hg) / r3 - c3 * r3_g * r3_h / den(r3 * r3);

    DS_PHASE_INNER();
    double b4 = 2.0 - d * d / e;
    double b4_h = e_h * dsd::ipow<2>(d / e) - 2.0 * d * d_h / e;
    double b4_hh = e_hh * dsd::ipow<2>(d / e) + 2.0 * e_h * dsd::ipow<2>(d / e) * (d_h / den(d) - e_h / e) -
                   2.0 * (d_h * d_h + d * d_hh) / e + 2. * d * d_h * e_h / e2D;

    double c4 = 2.0 * e * sqge + d;
    double c4_h = 2.0 * e_h * sqge + d_h;
    double c4_hh = 2.0 * e_hh * sqge + d_hh;
    double c4_g = e / sqge;
    double c4_gg = -0.5 * e / den(Gamma_e * sqge);
    double c4_hg = e_h / sqge;

    DS_PHASE_INNER();
    double s4a = 2.0 / e / discr;
    double s4ah = e_h / e + di_h / discr;
    double s4a_h = -s4a * s4ah;
    double s4a_hh = -s4a_h * s4ah - s4a * (e_hh / e - dsd::ipow<2>(e_h / e) + di_hh / discr - dsd::ipow<2>(di_h / discr));
    double s4b = d * b3 + b4 * b2;
    double s4b_h = d_h * b3 + d * b3_h + b4_h * b2 + b4 * b2_h;
    double s4b_hh = d_hh * b3 + 2.0 * d_h * b3_h + d * b3_hh + b4_hh * b2 + 2. * b4_h * b2_h + b4 * b2_hh;
    double s4c = dsm::atan(c4 / discr) - dsm::atan(d / discr);
    double up1 = c4_h * discr - c4 * di_h;
    const DsDen<FAST> dn1 = den(discr * discr + c4 * c4);
    double up2 = d_h * discr - d * di_h;
    double dn2 = discr * discr + d * d;
    double s4c_h = up1 / dn1 - up2 / den(dn2);
    double s4c_hh = (c4_hh * discr - c4 * di_hh) / dn1 - up1 * 2.0 * (discr * di_h + c4 * c4_h) / den(dn1 * dn1) -
                    (d_hh * discr - d * di_hh) / den(dn2) + up2 * 2.0 * (discr * di_h + d * d_h) / den(dn2 * dn2);
    double s4c_g = c4_g * discr / dn1;
    double s4c_gg = c4_gg * discr / dn1 - 2.0 * c4 * discr * dsd::ipow<2>(c4_g / dn1);
    double s4c_hg = (c4_hg * discr + c4_g * di_h - c4_g * discr / dn1 * 2. * (discr * di_h + c4 * c4_h)) / dn1;
    DS_PHASE_INNER();
    double s4 = s4a * s4b * s4c;
    double s4_h = s4a_h * s4b * s4c + s4a * s4b_h * s4c + s4a * s4b * s4c_h;
    double s4_hh = s4a_hh * s4b * s4c + s4a * s4b_hh * s4c + s4a * s4b * s4c_hh +
                   2.0 * (s4a_h * s4b_h * s4c + s4a_h * s4b * s4c_h + s4a * s4b_h * s4c_h);
    double s4_g = s4a * s4b * s4c_g;
    double s4_gg = s4a * s4b * s4c_gg;
    double s4_hg = s4a * s4b * s4c_hg + s4c_g * (s4a_h * s4b + s4a * s4b_h);

    double f = s1 + s2 + s3 + s4;
    double f_h = s1_h + s2_h + s3_h + s4_h;
    double f_g = s1_g + s2_g + s3_g + s4_g;
    double f_hh = s1_hh + s2_hh + s3_hh + s4_hh;
    double f_gg = s2_gg + s3_gg + s4_gg;
    double f_hg = s1_hg + s2_hg + s3_hg + s4_hg;

    double p = onethird * (Gamma_e * f_g - 2.0 * theta * f_h);
    double u = Gamma_e * f_g - theta * f_h;
    double s = (Gamma_e * s2_g - s2 + Gamma_e * s3_g - s3 + s4a * s4b * (Gamma_e * s4c_g - s4c)) - theta * f_h;
    if (fabs(s) < DS_R4(1.0e-9) * fabs(theta * f_h)) s = 0.0;
    double cv = 2.0 * theta * (Gamma_e * f_hg - f_h) - theta * theta * f_hh - Gamma_e * Gamma_e * f_gg;
    if (fabs(cv) < DS_R4(1.0e-9) * fabs(Gamma_e * Gamma_e * f_gg)) cv = 0.0;
    double pdlh = onethird * theta * (Gamma_e * f_hg - 2.0 * f_h - 2.0 * theta * f_hh);
    double pdlg = onethird * Gamma_e * (f_g + Gamma_e * f_gg - 2.0 * theta * f_hg);
    double dpr = p + onethird * (pdlg - 2.0 * pdlh);
    double dpt = Gamma_e * (theta * f_hg - onethird * Gamma_e * f_gg) - theta * (f_h / den(0.75) + theta * f_hh / den(1.5));
    r = DsThermo{f, u, p, s, cv, dpr, dpt};
}

template <bool FAST>
DS_D void ie_liquid(double Gamma_e, double rs, double Z, const DsIonZ& zc, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double cDH = zc.cDH, cTF = zc.cTF;   // Z-only pieces: DsIonZ
    double xrs = ds_hc.xrs;
    const DsDen<FAST> rsD = den(rs), Gamma_eD = den(Gamma_e);
    const DsDen<FAST> xr = den(xrs / rsD);
    const double a = zc.aL, b = zc.bL, nu = zc.nuL;
    double sGam = sq(Gamma_e);
    double nuGam = dsm::pow(Gamma_e, nu);
    double nuGam_g = nu * nuGam / Gamma_eD;
    double nuGam_gg = (nu - 1.0) * nuGam_g / Gamma_eD;
    double ty1 = 1.0 / den(DS_R4(0.001) * (Z * Z) + 2.0 * Gamma_e);
    double ty1_g = -2.0 * (ty1 * ty1);
    double ty1_gg = -4.0 * ty1 * ty1_g;
    double rs2 = rs * rs, rs3 = rs * rs * rs;
    const DsDen<FAST> ty2 = den(1.0 + 6.0 * rs2);
    double ty2_x = -12.0 * rs2 / xr;
    double ty2_xx = -3.0 * ty2_x / xr;
    double ty = rs3 / ty2 * (1.0 + ty1);
    double tyx = 3.0 / xr + ty2_x / ty2;
    double ty_x = -ty * tyx;
    double ty_g = rs3 * ty1_g / ty2;
    double p1 = (Z - 1.0) / den(9.0);
    double cor1 = 1.0 + p1 * ty;
    double cor1_x = p1 * ty_x;
    double cor1_g = p1 * ty_g;
    double cor1_xx = p1 * (ty * (3.0 / den(xr * xr) + dsd::ipow<2>(ty2_x / ty2) - ty2_xx / ty2) - ty_x * tyx);
    double cor1_gg = p1 * rs3 * ty1_gg / ty2;
    double cor1_xg = -p1 * ty_g * tyx;
    double u0 = DS_R4(0.78) * sq(Gamma_e / den(Z)) * rs3;
    double u0_x = -3.0 * u0 / xr;
    double u0_g = 0.5 * u0 / Gamma_eD;
    double u0_xx = -4.0 * u0_x / xr;
    double u0_gg = -0.5 * u0_g / Gamma_eD;
    double u0_xg = -3.0 * u0_g / xr;
    double d0_g = Z * Z * Z;
    const DsDen<FAST> d0 = den(Gamma_e * d0_g + 21.0 * rs3);
    double d0_x = -63.0 * rs3 / xr;
    double d0_xx = 252.0 * rs3 / den(xr * xr);
    double cor0 = 1.0 + u0 / d0;
    double cor0_x = (u0_x - u0 * d0_x / d0) / d0;
    double cor0_g = (u0_g - u0 * d0_g / d0) / d0;
    double cor0_xx = (u0_xx - (2.0 * u0_x * d0_x + u0 * d0_xx) / d0 + 2.0 * dsd::ipow<2>(d0_x / d0)) / d0;
    double cor0_gg = (u0_gg - 2.0 * u0_g * d0_g / d0 + 2.0 * u0 * dsd::ipow<2>(d0_g / d0)) / d0;
    double cor0_xg = (u0_xg - (u0_x * d0_g + u0_g * d0_x) / d0 + 2.0 * u0 * d0_x * d0_g / den(d0 * d0)) / d0;

    const DsDen<FAST> rgam = den(sq(1.0 + xr * xr));
    const double q1 = zc.q1L, q2 = zc.q2L;
    const DsDen<FAST> h1u = den(1.0 + DS_R4(0.2) * (xr * xr));
    const DsDen<FAST> h1d = den(1.0 + q1 * xr + q2 * (xr * xr));
    double h1 = h1u / h1d;
    double h1x = DS_R4(0.4) * xr / h1u - (q1 + 2. * q2 * xr) / h1d;
    double h1_x = h1 * h1x;
    double h1_xx = h1_x * h1x + h1 * (DS_R4(.4) / h1u - dsd::ipow<2>(DS_R4(.4) * xr / h1u) - 2. * q2 / h1d +
                                      dsd::ipow<2>((q1 + 2. * q2 * xr) / h1d));
    double up = cDH * sGam + a * cTF * nuGam * cor0 * h1;
    double up_x = a * cTF * nuGam * (cor0_x * h1 + cor0 * h1_x);
    double up_g = 0.5 * cDH / den(sGam) + a * cTF * (nuGam_g * cor0 + nuGam * cor0_g) * h1;
    double up_xx = a * cTF * nuGam * (cor0_xx * h1 + 2. * cor0_x * h1_x + cor0 * h1_xx);
    double up_gg = -0.25 * cDH / den(sGam * Gamma_e) +
                   a * cTF * (nuGam_gg * cor0 + 2.0 * nuGam_g * cor0_g + nuGam * cor0_gg) * h1;
    double up_xg = a * cTF * (nuGam_g * (cor0_x * h1 + cor0 * h1_x) + nuGam * (cor0_xg * h1 + cor0_g * h1_x));
    double dn1 = b * sGam + a / rsD * nuGam * cor1;
    double dn1_x = a * nuGam * (cor1 / den(xrs) + cor1_x / rsD);
    double dn1_g = 0.5 * b / den(sGam) + a / rsD * (nuGam_g * cor1 + nuGam * cor1_g);
    double dn1_xx = a * nuGam / den(xrs) * (2. * cor1_x + xr * cor1_xx);
    double dn1_gg = -0.25 * b / den(Gamma_e * sGam) +
                    a / rsD * (nuGam_gg * cor1 + 2. * nuGam_g * cor1_g + nuGam * cor1_gg);
    double dn1_xg = a * (nuGam_g * (cor1 / den(xrs) + cor1_x / rsD) + nuGam * (cor1_g / den(xrs) + cor1_xg / rsD));
    const DsDen<FAST> dn = den(1.0 + dn1 / rgam);
    double dn_x = dn1_x / rgam - xr * dn1 / den(dsd::ipow<3>(rgam));
    double dn_xx = (dn1_xx - ((2. * xr * dn1_x + dn1) - 3. * (xr * xr) * dn1 / den(rgam * rgam)) / den(rgam * rgam)) / rgam;
    double dn_g = dn1_g / rgam;
    double dn_gg = dn1_gg / rgam;
    double dn_xg = dn1_xg / rgam - xr * dn1_g / den(dsd::ipow<3>(rgam));
    double f = -up / dn * Gamma_e;
    double fx = (up * dn_x / dn - up_x) / dn;
    double fx_g = ((up_g * dn_x + up_x * dn_g + up * dn_xg - 2.0 * up * dn_x * dn_g / dn) / dn - up_xg) / dn;
    double f_x = fx * Gamma_e;
    double fg = (up * dn_g / dn - up_g) / dn;
    double f_g = fg * Gamma_e - up / dn;
    double f_xx = ((up * dn_xx + 2.0 * (up_x * dn_x - up * (dn_x * dn_x) / dn)) / dn - up_xx) / dn * Gamma_e;
    double f_gg = 2.0 * fg + Gamma_e * ((2. * dn_g * (up_g - up * dn_g / dn) + up * dn_gg) / dn - up_gg) / dn;
    double f_xg = fx + Gamma_e * fx_g;
    double u = Gamma_e * f_g;
    double s = u - f;
    double cv = -(Gamma_e * Gamma_e) * f_gg;
    double p = onethird * (xr * f_x + Gamma_e * f_g);
    double dpt = -(Gamma_e * Gamma_e) * (xr * fx_g + f_gg) * onethird;
    double dpr = (12.0 * p + (xr * xr) * f_xx + 2.0 * xr * Gamma_e * f_xg + (Gamma_e * Gamma_e) * f_gg) / den(9.0);
    r = DsThermo{f, u, p, s, cv, dpr, dpt};
}

template <bool FAST>
DS_D void ie_lattice(double Gamma, double theta, double rs, double Z, const DsIonZ& zc, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double a[4] = {DS_R4(1.1866), DS_R4(0.684), DS_R4(17.9), DS_R4(41.5)};
    const double px = DS_R4(0.205);
    double e1 = ds_hc.e1;
    double aTF = ds_hc.atf;
    double xrs = ds_hc.xrs;
    const DsDen<FAST> xr = den(xrs / den(rs));
    const double r3 = zc.r3;
    const double b[4] = {zc.b0, zc.b1, zc.b2, zc.b3};   // Z-only pieces: DsIonZ
    double finf = aTF * zc.Z23 * b[0] * sq(1.0 + b[1] / den(xr * xr));
    double finfx = -b[1] / den((b[1] + xr * xr) * xr);
    double finf_x = finf * finfx;
    double finf_xx = finf_x * finfx - finf_x * (b[1] + 3.0 * (xr * xr)) / den((b[1] + xr * xr) * xr);

    double q1u = b[2] + a[2] * (xr * xr);
    const DsDen<FAST> q1d = den(1.0 + b[3] * (xr * xr));
    double q1 = q1u / q1d;
    double q1x = 2.0 * xr * (a[2] / den(q1u) - b[3] / q1d);
    double q1x_x = q1x / xr + 4.0 * (xr * xr) * (dsd::ipow<2>(b[3] / q1d) - dsd::ipow<2>(a[2] / den(q1u)));
    double q1_x = q1 * q1x;
    double q1_xx = q1_x * q1x + q1 * q1x_x;

    double sup, sup_x, sup_g, sup_xx, sup_gg, sup_xg;
    if (theta < 6.0 / den(px)) {
        double y0 = dsd::ipow<2>(px * theta);
        double y0_x = y0 / xr;
        double y0_g = 2.0 * y0 / den(Gamma);
        double y0_xx = 0.0;
        double y0_gg = y0_g / den(Gamma);
        double y0_xg = y0_g / xr;
        const DsDen<FAST> y1 = den(dsm::exp(y0));
        double y1_x = y1 * y0_x;
        double y1_g = y1 * y0_g;
        double y1_xx = y1 * (y0_x * y0_x + y0_xx);
        double y1_gg = y1 * (y0_g * y0_g + y0_gg);
        double y1_xg = y1 * (y0_x * y0_g + y0_xg);
        const DsDen<FAST> sa = den(1.0 + y1);
        const DsDen<FAST> supa = den(dsm::log(sa));
        double supa_x = y1_x / sa;
        double supa_g = y1_g / sa;
        double supa_xx = (y1_xx - y1_x * y1_x / sa) / sa;
        double supa_gg = (y1_gg - y1_g * y1_g / sa) / sa;
        double supa_xg = (y1_xg - y1_x * y1_g / sa) / sa;
        double em2 = e1 - 2.0;
        double sb = e1 - em2 / y1;
        const DsDen<FAST> supb = den(dsm::log(sb));
        double em2y1 = em2 / den((y1 * y1) * sb);
        double supb_x = em2y1 * y1_x;
        double supb_g = em2y1 * y1_g;
        double supb_xx = em2y1 * (y1_xx - 2.0 * (y1_x * y1_x) / y1 - y1_x * supb_x);
        double supb_gg = em2y1 * (y1_gg - 2.0 * (y1_g * y1_g) / y1 - y1_g * supb_g);
        double supb_xg = em2y1 * (y1_xg - 2.0 * y1_x * y1_g / y1 - y1_g * supb_x);
        sup = sq(supa / supb);
        double supx = 0.5 * (supa_x / supa - supb_x / supb);
        sup_x = sup * supx;
        double supg = 0.5 * (supa_g / supa - supb_g / supb);
        sup_g = sup * supg;
        sup_xx = sup_x * supx + sup * 0.5 * (supa_xx / supa - dsd::ipow<2>(supa_x / supa) - supb_xx / supb +
                                             dsd::ipow<2>(supb_x / supb));
        sup_gg = sup_g * supg + sup * 0.5 * (supa_gg / supa - dsd::ipow<2>(supa_g / supa) - supb_gg / supb +
                                             dsd::ipow<2>(supb_g / supb));
        sup_xg = sup_x * supg + sup * 0.5 * ((supa_xg - supa_x * supa_g / supa) / supa -
                                             (supb_xg - supb_x * supb_g / supb) / supb);
    } else {
        sup = px * theta;
        sup_x = 0.5 * px * theta / xr;
        sup_g = px * theta / den(Gamma);
        sup_xx = -0.5 * sup_x / xr;
        sup_gg = 0.0;
        sup_xg = sup_x / den(Gamma);
    }
    const DsDen<FAST> supD = den(sup), GammaD = den(Gamma);
    const DsDen<FAST> gr3 = den(dsm::pow(Gamma / supD, r3));
    const DsDen<FAST> gr32D = den(gr3 * gr3);
    double gr3x = -r3 * sup_x / supD;
    double gr3_x = gr3 * gr3x;
    double gr3_xx = gr3_x * gr3x - r3 * gr3 * (sup_xx / supD - dsd::ipow<2>(sup_x / supD));
    double gr3g = r3 * (1.0 / GammaD - sup_g / supD);
    double gr3_g = gr3 * gr3g;
    double gr3_gg = gr3_g * gr3g + gr3 * r3 * (dsd::ipow<2>(sup_g / supD) - sup_gg / supD - 1.0 / den(Gamma * Gamma));
    double gr3_xg = gr3_g * gr3x + gr3 * r3 * (sup_x * sup_g / den(sup * sup) - sup_xg / supD);
    double w = 1.0 + q1 / gr3;
    double w_x = q1_x / gr3 - q1 * gr3_x / gr32D;
    double w_g = -q1 * gr3_g / gr32D;
    double w_xx = q1_xx / gr3 - (2.0 * q1_x * gr3_x + q1 * (gr3_xx - 2.0 * (gr3_x * gr3_x) / gr3)) / gr32D;
    double w_gg = q1 * (2.0 * (gr3_g * gr3_g) / gr3 - gr3_gg) / gr32D;
    double w_xg = -(q1_x * gr3_g + q1 * (gr3_xg - 2.0 * gr3_x * gr3_g / gr3)) / gr32D;

    double f = -Gamma * finf * w;
    double f_x = -Gamma * (finf_x * w + finf * w_x);
    double f_xx = -Gamma * (finf_xx * w + 2.0 * finf_x * w_x + finf * w_xx);
    double f_g = -finf * w - Gamma * finf * w_g;
    double f_gg = -2.0 * finf * w_g - Gamma * finf * w_gg;
    double f_xg = -finf_x * w - finf * w_x - Gamma * (finf_x * w_g + finf * w_xg);

    double s = -(Gamma * Gamma) * finf * w_g;
    double u = s + f;
    double cv = -(Gamma * Gamma) * f_gg;
    double p = onethird * (xr * f_x + Gamma * f_g);
    double dpt = onethird * (Gamma * Gamma) * (xr * finf * (finfx * w_g + w_xg) - f_gg);
    double dpr = (12.0 * p + (xr * xr) * f_xx + 2.0 * xr * Gamma * f_xg + (Gamma * Gamma) * f_gg) / den(9.0);
    r = DsThermo{f, u, p, s, cv, dpr, dpt};
}

template <bool FAST>
DS_D void ii_liquid(double Gamma, double theta, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    // array constructor of REAL(4): A(3) is evaluated entirely in single precision
    const float A1f = -0.907347f, A2f = 0.62849f;
    const float A3f = -0.5f * sqrtf(3.0f) + 0.907347f / sqrtf(0.62849f);
    const double A1 = (double)A1f, A2 = (double)A2f, A3 = (double)A3f;
    const double B1 = DS_R4(4.50e-3), B2 = DS_R4(170.0), B3 = DS_R4(-8.4e-5), B4 = DS_R4(3.70e-3);

    double f0 = A1 * (sq(Gamma * (A2 + Gamma)) -
                      A2 * dsm::log(sq(Gamma / den(A2)) + sq(1.0 + Gamma / den(A2)))) +
                2.0 * A3 * (sq(Gamma) - dsm::atan(sq(Gamma)));
    double G2 = Gamma * Gamma;
    double G15 = dsm::pow(Gamma, 1.5);
    double f = f0 + B1 * (Gamma - B2 * dsm::log(1.0 + Gamma / den(B2))) + 0.5 * B3 * dsm::log(1.0 + G2 / den(B4));
    double u = G15 * (A1 / den(sq(A2 + Gamma)) + A3 / den(1.0 + Gamma)) +
               G2 * (B1 / den(B2 + Gamma) + B3 / den(B4 + G2));
    double cv = 0.5 * G15 * (A3 * (Gamma - 1.0) / den(dsd::ipow<2>(Gamma + 1.0)) - A1 * A2 / den(dsm::pow(Gamma + A2, 1.5))) +
                G2 * (B3 * (G2 - B4) / den(dsd::ipow<2>(G2 + B4)) - B1 * B2 / den(dsd::ipow<2>(Gamma + B2)));
    double p = onethird * u;
    double dpr = (4.0 * u - cv) / den_c(9.0);
    double dpt = onethird * cv;

    double fid = 3.0 * dsm::log(theta) - 1.5 * dsm::log(Gamma) - ds_hc.half_log_6_over_pi - 1.0;
    f = f + fid;
    u = u + 1.5;
    p = p + 1.0;
    cv = cv + 1.5;
    double s = u - f;
    dpt = dpt + 1.0;
    dpr = dpr + 1.0;

    double fq = theta * theta / den_c(24.0);
    f = f + fq;
    u = u + 2.0 * fq;
    s = s + fq;
    cv = cv - 2.0 * fq;
    p = p + fq;
    dpr = dpr + 2.0 * fq;
    dpt = dpt - fq;
    r = DsThermo{f, u, p, s, cv, dpr, dpt};
}

template <bool FAST>
DS_D void ii_lattice(double Gamma, double theta, DsThermo& tot, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double KM = DS_R4(-0.895929255682), w = DS_R4(0.5113875), clm = DS_R4(-2.49389);
    const double alpha[9] = {DS_R4(0.0), DS_R4(0.932446), DS_R4(0.334547), DS_R4(0.265764), DS_R4(0.0),
                             DS_R4(0.0), DS_R4(4.757014e-3), DS_R4(0.0), DS_R4(4.7770935e-3)};
    const double a[9] = {DS_R4(1.0), DS_R4(0.1839), DS_R4(0.593586), DS_R4(5.4814e-3), DS_R4(5.01813e-4),
                         DS_R4(0.0), DS_R4(3.9247e-7), DS_R4(0.0), DS_R4(5.8356e-11)};
    const double b[8] = {DS_R4(261.66), DS_R4(0.0), DS_R4(7.07997), DS_R4(0.0), DS_R4(0.0409484),
                         DS_R4(3.97355e-4), DS_R4(5.11148e-5), DS_R4(2.19749e-6)};
    const double afh[4] = {0.0, DS_R4(10.9), DS_R4(247.0), DS_R4(1.765e5)};  // 1-based
    const double b1 = DS_R4(0.12);
    const double c1 = b1 / den(afh[1]);
    const double theta_low = DS_R4(1.0e-5), theta_high = DS_R4(1.0e5);

    double fh = 0, uh = 0, cvh = 0;
    if (theta > theta_high) {
        uh = 3.0 / den(alpha[6] * dsd::ipow<3>(theta));
        fh = -uh / den_c(3.0);
        cvh = 4.0 * uh;
    } else if (theta < theta_low) {
        fh = 3.0 * dsm::log(theta) + clm - 1.5 * w * theta + theta * theta / den_c(24.0);
        uh = 3.0 - 1.5 * w * theta + theta * theta / den_c(12.0);
        cvh = 3.0 - theta * theta / den_c(12.0);
    } else {
        // Ath / Bth : value, first and second derivative by Horner (:545-575)
        double Ac[3] = {a[8], 0.0, 0.0};
        for (int n = 7; n >= 0; --n) {
            Ac[2] = Ac[2] * theta + Ac[1];
            Ac[1] = Ac[1] * theta + Ac[0];
            Ac[0] = Ac[0] * theta + a[n];
        }
        Ac[2] = Ac[2] * 2.0;
        double Bc[3] = {b[7], 0.0, 0.0};
        for (int n = 6; n >= 0; --n) {
            Bc[2] = Bc[2] * theta + Bc[1];
            Bc[1] = Bc[1] * theta + Bc[0];
            Bc[0] = Bc[0] * theta + b[n];
        }
        Bc[2] = Bc[2] * 2.0;
        Bc[0] = Bc[0] + alpha[6] * a[6] * dsd::ipow<9>(theta) + alpha[8] * a[8] * dsd::ipow<11>(theta);
        Bc[1] = Bc[1] + 9.0 * alpha[6] * a[6] * dsd::ipow<8>(theta) + 11.0 * alpha[8] * a[8] * dsd::ipow<10>(theta);
        Bc[2] = Bc[2] + 72.0 * alpha[6] * a[6] * dsd::ipow<7>(theta) + 110.0 * alpha[8] * a[8] * dsd::ipow<9>(theta);
        for (int n = 1; n <= 3; ++n) {
            double ea = dsm::exp(-alpha[n] * theta);
            double omea = 1.0 - ea;
            fh = fh + dsm::log(omea);
            uh = uh + alpha[n] * ea / omea;  // ea = exp(-alpha theta) reaches the subnormals: plain divisions
            cvh = cvh + alpha[n] * alpha[n] * ea / (omea * omea);
        }
        fh = fh - Ac[0] / den(Bc[0]);
        uh = (uh - (Ac[1] * Bc[0] - Ac[0] * Bc[1]) / den(Bc[0] * Bc[0])) * theta;
        cvh = cvh + ((Ac[2] * Bc[0] - Bc[2] * Ac[0]) * Bc[0] - 2.0 * (Ac[1] * Bc[0] - Bc[1] * Ac[0]) * Bc[1]) /
                        dsd::ipow<3>(Bc[0]);
        cvh = cvh * (theta * theta);
    }
    double e0 = 1.5 * w * theta;
    double sh = uh - fh;
    uh = uh + e0;
    fh = fh + e0;
    double ph = 0.5 * uh;
    double dpth = 0.5 * cvh;
    double dprh = 0.75 * uh - 0.25 * cvh;

    // anharmonic
    double fa = 0, ua = 0, pa = 0, cva = 0, dpta = 0, dpra = 0;
    double theta2 = theta * theta;
    double theta4 = dsd::ipow<4>(theta);
    double c1t2 = c1 * theta2;
    double sup = dsm::exp(-c1t2);
    double tq = b1 * theta2 / den(Gamma);
    double gninv = sup;
    for (int n = 1; n <= 3; ++n) {
        gninv = gninv / Gamma;  // exp(-c1 theta^2) / Gamma^n underflows gracefully: plain division
        double ninv = (double)(1.0f / (float)n);  // REAL(4) 1.0/n
        fa = fa - ninv * afh[n] * gninv;
        ua = ua + gninv * afh[n] * (1.0 + 2.0 * ninv * c1t2);
        double pfac = onethird * afh[n] + c1t2 * afh[n] * ninv;
        pa = pa + pfac * gninv;
        cva = cva + gninv * ((1.0 + n) * afh[n] + (4.0 - 2.0 * ninv) * afh[n] * c1t2 +
                             4.0 * afh[n] * (c1 * c1) * ninv * theta4);
        dpta = dpta + (pfac * (1.0 + n + 2.0 * c1t2) - 2.0 * ninv * afh[n] * c1t2) * gninv;
        dpra = dpra + (pfac * (1.0 - onethird * n - c1t2) + afh[n] * ninv * c1t2) * gninv;
    }
    fa = fa - tq;
    ua = ua - tq;
    pa = pa - 2.0 * onethird * tq;
    dpra = dpra - tq / den_c(4.5);

    double madelung = KM * Gamma;
    tot.f = fh + fa + madelung;
    tot.u = uh + ua + madelung;
    tot.p = ph + pa + onethird * madelung;
    tot.s = sh;
    tot.cv = cvh + cva;
    tot.dpt = dpth + dpta;
    tot.dpr = dprh + dpra + madelung / den_c(2.25);
}

template <bool FAST>
DS_D void one_ion(double Gamma_e, double rs, double Z, double A, const DsIonZ& zc, int phase, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    double Gamma = Gamma_e * zc.Z53;
    double theta = Gamma * sq(3.0 * Melectron / den(A) / den_c(amu) / den(rs)) / den(zc.Z76);
    DsThermo ie, ii;
    if (phase == ds_liquid_phase) {
        ie_liquid<FAST>(Gamma_e, rs, Z, zc, ie, bad);
        ii_liquid<FAST>(Gamma, theta, ii, bad);
    } else {
        ie_lattice<FAST>(Gamma, theta, rs, Z, zc, ie, bad);
        ii_lattice<FAST>(Gamma, theta, ii, bad);
    }
    r.f = ie.f + ii.f; r.u = ie.u + ii.u; r.p = ie.p + ii.p; r.s = ie.s + ii.s;
    r.cv = ie.cv + ii.cv; r.dpr = ie.dpr + ii.dpr; r.dpt = ie.dpt + ii.dpt;
}

template <bool FAST>
DS_D void LM_corrections(double rs, double Gamma_e, const DsComp& c, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double threshold = DS_R4(1.0e-9);
    double Gamma = Gamma_e * c.Z53;
    double dif0;
    if (rs < threshold) dif0 = c.Z52 - sq(dsd::ipow<3>(c.Z2) / den(c.Z));
    else dif0 = c.ZZ1_32 - sq(dsd::ipow<3>(c.Z2 + c.Z) / den(c.Z));
    double difR = dif0 / den(c.Z52);
    double difFDH = dif0 * Gamma_e * sq(onethird * Gamma_e);
    double D = c.Z2 / den(c.Z * c.Z);
    r = DsThermo{0, 0, 0, 0, 0, 0, 0};
    if (fabs(D - 1.0) < threshold) return;
    double p3 = dsm::pow(D, DS_R4(-0.2));
    double d0 = (DS_R4(2.6) * difR + 14.0 * dsd::ipow<3>(difR)) / den(1.0 - p3);
    double gp = d0 * dsm::pow(Gamma, p3);
    double f0 = difFDH / den(1.0 + gp);
    double q = (D * D) * DS_R4(0.0117);
    double rr = 1.5 / den(p3) - 1.0;
    double gq = q * gp;
    double f = f0 / den(dsm::pow(1.0 + gq, rr));
    double g = 1.5 - p3 * gp / den(1.0 + gp) - rr * p3 * gq / den(1.0 + gq);
    double u = f * g;
    double s = u - f;
    double p = onethird * u;
    double g_dg = -(p3 * p3) * (gp / den(dsd::ipow<2>(1.0 + gp)) + rr * gq / den(dsd::ipow<2>(1.0 + gq)));
    double u_dg = u * g + f * g_dg;
    double cv = u - u_dg;
    double dpr = p + u_dg / den_c(9.0);
    double dpt = p - onethird * u_dg;
    r = DsThermo{f, u, p, s, cv, dpr, dpt};
}

template <bool FAST>
DS_D int ion_mixture(const DsEosCtrl& rq, double rs, double Gamma_e, const DsComp& c,
                int ncharged, const double* Zion, const double* Aion, const DsIonZ* __restrict__ ionz,
                const double* Yion, double& Gamma, double& ionQ, int& phase, DsThermo& r, bool& bad,
                const unsigned char* act = nullptr, int nact = 0) {
    DS_DEN_SCOPE(FAST);
    const double Q2_threshold = 18.0;
    int err = 0;
    double Mu = amu / den_c(Melectron);
    Gamma = Gamma_e * c.Z53;
    phase = (Gamma < rq.Gamma_melt) ? ds_liquid_phase : ds_solid_phase;
    double ionQ2 = 3.0 * (Gamma_e * Gamma_e) / den(rs) / den(Mu) * c.Z2XoA2 * c.A / den(c.Z);
    ionQ = sq(ionQ2);
    if (phase == ds_liquid_phase && ionQ2 > Q2_threshold) err = 1;

    double f = 0, u = 0, p = 0, s = 0, cv = 0, dpr = 0, dpt = 0;
    // `act` (optional): the indices of the species above the abundance threshold, ascending -- the coefficient pass
    // builds it once per zone, so the loop (whose head reloads spilled registers) runs 1-2 times instead of ncharged
    const int n_iter = act ? nact : ncharged;
    for (int it_ = 0; it_ < n_iter; ++it_) {
        const int i = act ? act[it_] : it_;
        if (Yion[i] > rq.Ythresh) {
            DsThermo t;
            one_ion<FAST>(Gamma_e, rs, Zion[i], Aion[i], ionz[i], phase, t, bad);
            f = f + Yion[i] * t.f;
            u = u + Yion[i] * t.u;
            s = s + Yion[i] * (t.s - dsm::log(Yion[i]));
            p = p + Yion[i] * t.p;
            cv = cv + Yion[i] * t.cv;
            dpr = dpr + Yion[i] * t.dpr;
            dpt = dpt + Yion[i] * t.dpt;
        }
    }
    DsThermo m;
    LM_corrections<FAST>(rs, Gamma_e, c, m, bad);
    r.f = f * c.A + m.f;
    r.u = u * c.A + m.u;
    r.s = s * c.A + m.s;
    r.p = p * c.A + m.p;
    r.cv = cv * c.A + m.cv;
    r.dpr = dpr * c.A + m.dpr;
    r.dpt = dpt * c.A + m.dpt;
    return err;
}

DS_D double ratfit(const double* a, int m, const double* b, int k, double xx) {
    // rn = xx + a(m); rn = rn*xx + a(i), i=m-1..1 ; den = b(k+1); den = den*xx + b(i), i=k..1
    double rn = xx + a[m - 1];
    for (int i = m - 2; i >= 0; --i) rn = rn * xx + a[i];
    double den = b[k];
    for (int i = k - 1; i >= 0; --i) den = den * xx + b[i];
    return rn / den;
}

DS_D double zfermim12(double x) {  // funct_fermi1.f:21-86
    if (x < 2.0) { double xx = dsm::exp(x); return xx * ratfit(zfermim12_a1, 7, zfermim12_b1, 7, xx); }
    double xx = 1.0 / (x * x);
    return sqrt(x) * ratfit(zfermim12_a2, 11, zfermim12_b2, 11, xx);
}

DS_D double zfermi12(double x) {  // :88-146
    if (x < 2.0) { double xx = dsm::exp(x); return xx * ratfit(zfermi12_a1, 7, zfermi12_b1, 7, xx); }
    double xx = 1.0 / (x * x);
    return x * sqrt(x) * ratfit(zfermi12_a2, 10, zfermi12_b2, 11, xx);
}

DS_D double zfermi32(double x) {  // :216-276
    if (x < 2.0) { double xx = dsm::exp(x); return xx * ratfit(zfermi32_a1, 6, zfermi32_b1, 7, xx); }
    double xx = 1.0 / (x * x);
    return x * x * sqrt(x) * ratfit(zfermi32_a2, 9, zfermi32_b2, 10, xx);
}

DS_D double ifermi12(double f) {  // :528-580
    const double an = 0.5;
    if (f < 4.0) return dsm::log(f * ratfit(ifermi12_a1, 4, ifermi12_b1, 3, f));
    double ff = 1.0 / dsm::pow(f, 1.0 / (1.0 + an));
    double rn = ff + ifermi12_a2[5];
    for (int i = 4; i >= 0; --i) rn = rn * ff + ifermi12_a2[i];
    double den = ifermi12_b2[5];
    for (int i = 4; i >= 0; --i) den = den * ff + ifermi12_b2[i];
    return rn / (den * ff);
}

template <bool FAST>
DS_D void MB77(double n, double T, double Tns, const DsZonePre& zp, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double c[4] = {DS_R4(1.2974), DS_R4(15.0298), DS_R4(-15.2343), DS_R4(7.4663)};
    r = DsThermo{0, 0, 0, 0, 0, 0, 0};
    if (n == 0.0) return;
    double k = zp.k_mb;   // (3 pi^2 n / 2)^(1/3) / cm_to_fm: zone-level, see DsZonePre
    double dkdn = onethird * k / den(n);
    double u = k * (c[0] + k * (c[1] + k * (c[2] + k * c[3])));
    u = u * mev_to_ergs;
    double p = onethird * n * k * (c[0] + k * (2.0 * c[1] + k * (3.0 * c[2] + k * 4.0 * c[3])));
    p = p * mev_to_ergs;
    double dpdk = onethird * n * k * (c[0] + k * (4.0 * c[1] + k * (9.0 * c[2] + 16.0 * c[3] * k))) * mev_to_ergs;
    double dpr = (p / den(n) + dpdk * dkdn) * n;
    double dpT = 0.0;

    double lambda3 = pi * pi * dsm::pow(hbar * hbar / den(Mneutron * boltzmann * T), 1.5) / den_c((double)sqrtf(2.0f));
    double zeta = ifermi12(n * lambda3);
    double cv = 2.5 * zfermi32(zeta) / den(zfermi12(zeta)) - 4.5 * zfermi12(zeta) / den(zfermim12(zeta));
    cv = cv * boltzmann;
    double s = fivethird * zfermi32(zeta) / den(zfermi12(zeta)) - zeta;
    s = s * boltzmann;
    double R;
    if (T < Tns) {
        double tau = T / den(Tns);
        double v = sq(1.0 - tau) * (DS_R4(1.456) - DS_R4(0.157) / den(sq(tau)) + DS_R4(1.764) / den(tau));
        R = dsm::pow(DS_R4(0.4186) + sq((double)(1.007f * 1.007f) + dsd::ipow<2>(DS_R4(0.5010) * v)), 2.5) *
            dsm::exp(DS_R4(1.456) - sq((double)(1.456f * 1.456f) + v * v));
    } else {
        R = 1.0;
    }
    cv = cv * R;
    s = s * R;
    double f = u - T * s;
    r = DsThermo{f, u, p, s, cv, dpr, dpT};
}

template <bool FAST>
DS_D void get_radiation_eos(double rho, double T, DsThermo& r, bool& bad) {
    DS_DEN_SCOPE(FAST);
    const double fourthird = 4.0 * onethird;
    double aT3 = arad * dsd::ipow<3>(T);
    double aT4 = arad * dsd::ipow<4>(T);
    r.u = aT4 / den(rho);
    r.p = onethird * aT4;
    r.f = -onethird * aT4 / den(rho);
    r.s = fourthird * aT3 / den(rho);
    r.dpr = 0.0;
    r.dpt = fourthird * aT4;
    r.cv = 4.0 * aT3 / den(rho);
}

}  // namespace dse
using dse::DsIonZ;
using dse::ds_ionz_compute;

// Result of one EOS evaluation; fields follow dStar_eos_def.f:6-36 (res(1:15)).
struct DsEosRes {
    double lnP, lnE, lnS, grad_ad, chiRho, chiT, Cp, Cv, Chi, Gamma, Theta, mu_e, mu_n, Gamma1, Gamma3;
    int phase, ierr;
};

// Sum of the five components (dStar_eos_lib.f:173-258).  `logT` is dsm::log10(T); for the coefficient
// pass it is supplied by the host together with T so that table-cell selection is identical on
// both sides of the ABI.  Zion/Aion/Yion: the charged species of the zone (Yion = Y/(1-Yn)).
// FAST selects the branch-free divisions (DsDen, dconst.cuh); `bad` is raised when an operand left their
// safe range, and the caller then repeats the evaluation with FAST = false (all lanes of the CTA together,
// because of the DS_PHASE barriers inside).
// COMP selects the components to evaluate (the others contribute zero): a kernel that holds only some of them writes
// partial sums that a later kernel completes (kernels_micro.cuh, split cell pass).
enum { DS_EOS_ELECTRON = 1, DS_EOS_EXCHANGE = 2, DS_EOS_ION = 4, DS_EOS_NEUTRON = 8, DS_EOS_RADIATION = 16, DS_EOS_ALL = 31 };
template <bool FAST, int COMP = DS_EOS_ALL>
DS_D void ds_eval_crust_eos(const DsHelmDev& helm, const DsEosCtrl& rq, double rho, double T,
                            double logT, const DsComp& c, const DsZonePre& zp, const DsHelmZone& hz, int ncharged,
                            const double* __restrict__ Zion, const double* __restrict__ Aion,
                            const DsIonZ* __restrict__ ionz, const double* __restrict__ Yion, double Tns, double chi,
                            DsEosRes& o, bool& bad,
                            double* __restrict__ comps = nullptr, const unsigned char* act = nullptr, int nact = 0) {
    DS_DEN_SCOPE(FAST);
    using namespace dsc;
    double ne, rs, Gamma_e, Gamma, ionQ;
    dse::lattice_properties<FAST>(rho, T, c, zp, ne, rs, Gamma_e, Gamma, ionQ, bad);
    DS_PHASE();

    // electrons: table + exchange (electron.f:10-51, :53-250)
    DsThermo e{0, 0, 0, 0, 0, 0, 0};
    double eta_e = 0.0;
    int ierr_h = 0;
    if ((COMP & DS_EOS_ELECTRON) && !(fabs(c.Yn - 1.0) <= (double)1.1920929e-07f)) {
        DsHelmOut he;
        ierr_h = ds_helm_electron(helm, T, logT, hz, he);   // hz: ds_helm_zone(helm, rho, A / (1 - Yn), Z)
        if (ierr_h == 0) {
            e.u = he.eele; e.p = he.pele; e.s = he.sele; e.f = e.u - T * e.s;
            e.cv = he.deept; e.dpr = rho * he.dpepd; e.dpt = T * he.dpept;
            eta_e = he.etaele;
        }
    }
    DS_PHASE();
    DsThermo ex{0, 0, 0, 0, 0, 0, 0};
    if (COMP & DS_EOS_EXCHANGE) {
        dse::ee_exchange<FAST>(Gamma_e, rs, ex, bad);
        DS_PHASE();
    }
    const double nek = ne * boltzmann;
    const double nekT = nek * T;
    const double uexfac = nekT / den(rho), pexfac = nekT, sexfac = nek / den(rho);

    // ions (ion.f:20-92)
    DsThermo ion{0, 0, 0, 0, 0, 0, 0};
    int phase = 0;
    if (COMP & DS_EOS_ION) {
        dse::ion_mixture<FAST>(rq, rs, Gamma_e, c, ncharged, Zion, Aion, ionz, Yion, Gamma, ionQ, phase, ion, bad, act, nact);
#ifndef DS_NO_PHASE_ION
        DS_PHASE();
#endif
    }
    const double n_i = ne / den(c.Z);
    const double nik = n_i * boltzmann;
    const double nikT = nik * T;
    const double uifac = nikT / den(rho), pifac = nikT, sifac = nik / den(rho);

    // dripped neutrons (neutron.f:15-69)
    const double nn = rho * c.Yn / den_c(amu) / den(1.0 - chi);
    DsThermo nt{0, 0, 0, 0, 0, 0, 0};
    if (COMP & DS_EOS_NEUTRON) {
        dse::MB77<FAST>(nn, T, Tns, zp, nt, bad);
        DS_PHASE();
    }
    const double unfac = avogadro * c.Yn, pnfac = 1.0, snfac = unfac;

    DsThermo rad{0, 0, 0, 0, 0, 0, 0};
    if (COMP & DS_EOS_RADIATION) dse::get_radiation_eos<FAST>(rho, T, rad, bad);

    const double p = e.p + ex.p * pexfac + ion.p * pifac + nt.p * pnfac + rad.p;
    const double u = e.u + ex.u * uexfac + ion.u * uifac + nt.u * unfac + rad.u;
    const double s = e.s + ex.s * sexfac + ion.s * sifac + nt.s * snfac + rad.s;
    const double cv = e.cv + ex.cv * sexfac + ion.cv * sifac + nt.cv * snfac + rad.cv;
    const double dpr = e.dpr + ex.dpr * pexfac + ion.dpr * pifac + nt.dpr * pnfac + rad.dpr;
    const double dpt = e.dpt + ex.dpt * pexfac + ion.dpt * pifac + nt.dpt * pnfac + rad.dpt;
    const double gamma3m1 = dpt / den(rho) / den(T) / den(cv);
    const double gamma1 = (dpr + dpt * gamma3m1) / den(p);
    DS_PHASE();
    o.lnP = dsm::log(p);
    o.lnE = dsm::log(u);
    o.lnS = dsm::log(s);
    o.grad_ad = gamma3m1 / den(gamma1);
    o.chiRho = dpr / den(p);
    o.chiT = dpt / den(p);
    o.Cp = gamma1 * p * cv / den(dpr);
    o.Cv = cv;
    o.Chi = chi;
    o.Gamma = Gamma;
    o.Theta = 1.0 / den(ionQ);
    o.mu_e = (eta_e * boltzmann * T) * ergs_to_mev;
    o.mu_n = (c.Yn < rq.Ythresh) ? 0.0 : (nt.u - T * nt.s + nt.p / nn) * ergs_to_mev;
    o.Gamma1 = gamma1;
    o.Gamma3 = gamma3m1;
    o.phase = phase;
    o.ierr = (chi < 1.0) ? ierr_h : -10;
    if (comps) {  // optional `components` argument of eval_crust_eos (dStar_eos_lib.f:260-296): (7,4) = P,E,S,F,Cv,dPdlnRho,dPdlnT
        const double pe = e.p + ex.p * pexfac;
        double* q = comps;
        q[0] = pe; q[1] = e.u + ex.u * uexfac; q[2] = e.s + ex.s * sexfac; q[3] = e.f + ex.f * uexfac;
        q[4] = e.cv + ex.cv * sexfac; q[5] = (e.dpr + ex.dpr * pexfac) / den(pe); q[6] = (e.dpt + ex.dpt * pexfac) / den(pe);
        q += 7;
        q[0] = ion.p * pifac; q[1] = ion.u * uifac; q[2] = ion.s * sifac; q[3] = ion.f * uifac; q[4] = ion.cv * sifac;
        q[5] = ion.dpr / den(ion.p); q[6] = ion.dpt / den(ion.p);
        q += 7;
        if (c.Yn < rq.Ythresh) { for (int i = 0; i < 7; ++i) q[i] = 0.0; }
        else {
            q[0] = nt.p * pnfac; q[1] = nt.u * unfac; q[2] = nt.s * snfac; q[3] = nt.f * unfac; q[4] = nt.cv * snfac;
            q[5] = nt.dpr / den(nt.p); q[6] = nt.dpt / den(nt.p);
        }
        q += 7;
        q[0] = rad.p; q[1] = rad.u; q[2] = rad.s; q[3] = rad.f; q[4] = rad.cv; q[5] = rad.dpr / den(rad.p); q[6] = rad.dpt / den(rad.p);
    }
}
