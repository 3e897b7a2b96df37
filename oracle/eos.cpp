// ORACLE (test infrastructure): crust equation of state.
// Restates dStar_eos/public/dStar_eos_lib.f:84-297, private/electron.f:53-250 (ee_exchange),
// private/ion.f:20-576, private/neutron.f:15-69, private/radiation.f:5-25,
// private/funct_fermi1.f (Antia fits).  REAL(4) literal semantics are kept via R4().
#include <algorithm>
#include <cmath>

#include "dstar_oracle.hpp"

namespace dso {
using namespace cst;

// dStar_eos_lib.f:84-98
void lattice_properties(double rho, double T, const Composition& c, double& ne, double& rs,
                        double& Gamma_e, double& Gamma, double& ionQ) {
    ne = rho * c.Ye / amu;
    rs = dsm::pow(3.0 / fourpi / ne, onethird) / a_Bohr;
    Gamma_e = 2.0 * Rydberg / boltzmann / T / rs;
    Gamma = Gamma_e * c.Z53;
    ionQ = std::sqrt(3.0 * (Gamma_e * Gamma_e) * Melectron / rs / amu * c.Z2XoA2 * c.A / c.Z);
}

// dStar_eos_lib.f:100-109
double nuclear_volume_fraction(double rho, const Composition& c, double nuclear_radius) {
    return onethird * fourpi * ipow(nuclear_radius * fm_to_cm, 3) * (rho * (1.0 - c.Yn) / amu);
}

// dStar_eos_lib.f:111-121
double neutron_wavenumber(double rho, const Composition& c, double chi) {
    double n = rho * c.Yn / amu / (1.0 - chi);
    return dsm::pow(threepisquare * n, onethird) / cm_to_fm;
}

// electron.f:53-250 -- Potekhin's electron exchange fit and derivatives
void ee_exchange(double Gamma_e, double rs, Thermo& r) {
    // theta_fac = 2.0*(4.0/9.0/pi)**(2.0/3.0): 4.0/9.0 and 2.0/3.0 are REAL(4) operations
    static const double theta_fac =
        2.0 * dsm::pow((double)(4.0f / 9.0f) / pi, (double)(2.0f / 3.0f));
    const double theta_low = R4(0.005);
    double theta = theta_fac * rs / Gamma_e;
    double sqth = std::sqrt(theta);
    double theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta3 * theta;
    double t1, t2, t1_h, t2_h, t1_hh, t2_hh;
    if (theta > theta_low) {
        double cht1 = dsm::cosh(1.0 / theta), sht1 = dsm::sinh(1.0 / theta);
        double cht2 = dsm::cosh(1.0 / sqth), sht2 = dsm::sinh(1.0 / sqth);
        t1 = sht1 / cht1;
        t2 = sht2 / cht2;
        t1_h = -1.0 / ipow(theta * cht1, 2);
        t1_hh = 2.0 / ipow(theta * cht1, 3) * (cht1 - sht1 / theta);
        t2_h = -0.5 * sqth / ipow(theta * cht2, 2);
        t2_hh = (0.75 * sqth * cht2 - 0.5 * sht2) / ipow(theta * cht2, 3);
    } else {
        t1 = 1.0; t2 = 1.0; t1_h = 0.0; t2_h = 0.0; t1_hh = 0.0; t2_hh = 0.0;
    }
    double a0 = 0.75 + R4(3.04363) * theta2 - R4(0.09227) * theta3 + R4(1.7035) * theta4;
    double a0_h = R4(6.08726) * theta - R4(0.27681) * theta2 + R4(6.814) * theta3;
    double a0_hh = R4(6.08726) - R4(0.55362) * theta + R4(20.442) * theta2;
    double a1 = 1. + R4(8.31051) * theta2 + R4(5.1105) * theta4;
    double a1_h = R4(16.62102) * theta + R4(20.442) * theta3;
    double a1_hh = R4(16.62102) + R4(61.326) * theta2;
    double a = R4(0.610887) * a0 / a1 * t1;
    double ah = a0_h / a0 - a1_h / a1 + t1_h / t1;
    double a_h = a * ah;
    double a_hh = a_h * ah + a * (a0_hh / a0 - ipow(a0_h / a0, 2) - a1_hh / a1 + ipow(a1_h / a1, 2) +
                                  t1_hh / t1 - ipow(t1_h / t1, 2));

    double b0 = R4(0.341308) + R4(12.070873) * theta2 + R4(1.148889) * theta4;
    double b0_h = R4(24.141746) * theta + R4(4.595556) * theta3;
    double b0_hh = R4(24.141746) + R4(13.786668) * theta2;
    double b1 = 1.0 + R4(10.495346) * theta2 + R4(1.326623) * theta4;
    double b1_h = R4(20.990692) * theta + R4(5.306492) * theta3;
    double b1_hh = R4(20.990692) + R4(15.919476) * theta2;
    double b = sqth * t2 * b0 / b1;
    double bh = 0.5 / theta + t2_h / t2 + b0_h / b0 - b1_h / b1;
    double b_h = b * bh;
    double b_hh = b_h * bh + b * (-0.5 / theta2 + t2_hh / t2 - ipow(t2_h / t2, 2) + b0_hh / b0 -
                                  ipow(b0_h / b0, 2) - b1_hh / b1 + ipow(b1_h / b1, 2));

    double d0 = R4(0.614925) + R4(16.996055) * theta2 + R4(1.489056) * theta4;
    double d0_h = R4(33.99211) * theta + R4(5.956224) * theta3;
    double d0_hh = R4(33.99211) + R4(17.868672) * theta2;
    double d1 = 1.0 + R4(10.10935) * theta2 + R4(1.22184) * theta4;
    double d1_h = R4(20.2187) * theta + R4(4.88736) * theta3;
    double d1_hh = R4(20.2187) + R4(14.66208) * theta2;
    double d = sqth * t2 * d0 / d1;
    double dh = 0.5 / theta + t2_h / t2 + d0_h / d0 - d1_h / d1;
    double d_h = d * dh;
    double d_hh = d_h * dh + d * (-0.5 / theta2 + t2_hh / t2 - ipow(t2_h / t2, 2) + d0_hh / d0 -
                                  ipow(d0_h / d0, 2) - d1_hh / d1 + ipow(d1_h / d1, 2));

    double e0 = R4(0.539409) + R4(2.522206) * theta2 + R4(0.178484) * theta4;
    double e0_h = R4(5.044412) * theta + R4(0.713936) * theta3;
    double e0_hh = R4(5.044412) + R4(2.141808) * theta2;
    double e1 = 1.0 + R4(2.555501) * theta2 + R4(0.146319) * theta4;
    double e1_h = R4(5.111002) * theta + R4(0.585276) * theta3;
    double e1_hh = R4(5.111002) + R4(1.755828) * theta2;
    double e = theta * t1 * e0 / e1;
    double eh = 1.0 / theta + t1_h / t1 + e0_h / e0 - e1_h / e1;
    double e_h = e * eh;
    double e_hh = e_h * eh + e * (t1_hh / t1 - ipow(t1_h / t1, 2) + e0_hh / e0 - ipow(e0_h / e0, 2) -
                                  e1_hh / e1 + ipow(e1_h / e1, 2) - 1.0 / theta2);

    double exp1th = dsm::exp(-1.0 / theta);
    double c = (R4(0.872496) + R4(0.025248) * exp1th) * e;
    double c_h = R4(0.025248) * exp1th / theta2 * e + c * e_h / e;
    double c_hh = R4(0.025248) * exp1th / theta2 * (e_h + (1.0 - 2.0 * theta) / theta2 * e) +
                  c_h * e_h / e + c * e_hh / e - c * ipow(e_h / e, 2);
    double discr = std::sqrt(4.0 * e - d * d);
    double di_h = 0.5 / discr * (4.0 * e_h - 2.0 * d * d_h);
    double di_hh = (-ipow((2.0 * e_h - d * d_h) / discr, 2) + 2.0 * e_hh - d_h * d_h - d * d_hh) / discr;

    double s1 = -c / e * Gamma_e;
    double s1h = c_h / c - e_h / e;
    double s1_h = s1 * s1h;
    double s1_hh = s1_h * s1h + s1 * (c_hh / c - ipow(c_h / c, 2) - e_hh / e + ipow(e_h / e, 2));
    double s1_g = -c / e;
    double s1_hg = s1_g * (c_h / c - e_h / e);
    double b2 = b - c * d / e;
    double b2_h = b_h - (c_h * d + c * d_h) / e + c * d * e_h / (e * e);
    double b2_hh = b_hh - (c_hh * d + 2.0 * c_h * d_h + c * d_hh) / e +
                   (2.0 * (c_h * d + c * d_h - c * d * e_h / e) * e_h + c * d * e_hh) / (e * e);

    double sqge = std::sqrt(Gamma_e);
    double s2 = -2.0 / e * b2 * sqge;
    double s2h = b2_h / b2 - e_h / e;
    double s2_h = s2 * s2h;
    double s2_hh = s2_h * s2h + s2 * (b2_hh / b2 - ipow(b2_h / b2, 2) - e_hh / e + ipow(e_h / e, 2));
    double s2_g = 0.5 * s2 / Gamma_e;
    double s2_gg = -0.5 * s2_g / Gamma_e;
    double s2_hg = 0.5 * s2_h / Gamma_e;

    double r3 = e * Gamma_e + d * sqge + 1.0;
    double r3_h = e_h * Gamma_e + d_h * sqge;
    double r3_hh = e_hh * Gamma_e + d_hh * sqge;
    double r3_g = e + 0.5 * d / sqge;
    double r3_gg = -0.25 * d / (Gamma_e * sqge);
    double r3_hg = e_h + 0.5 * d_h / sqge;

    double b3 = a - c / e;
    double b3_h = a_h - c_h / e + c * e_h / (e * e);
    double b3_hh = a_hh - c_hh / e + (2.0 * c_h * e_h + c * e_hh) / (e * e) - 2.0 * c * (e_h * e_h) / ipow(e, 3);

    double c3 = (d / e * b2 - b3) / e;
    double c3_h = (d_h * b2 + d * b2_h + b3 * e_h) / (e * e) - 2.0 * d * b2 * e_h / ipow(e, 3) - b3_h / e;
    double c3_hh = (-b3_hh + (d_hh * b2 + 2.0 * d_h * b2_h + d * b2_hh + b3_h * e_h + b3 * e_hh + b3_h * e_h) / e -
                    2.0 * ((d_h * b2 + d * b2_h + b3 * e_h + d_h * b2 + d * b2_h) * e_h + d * b2 * e_hh) / (e * e) +
                    6.0 * d * b2 * (e_h * e_h) / ipow(e, 3)) / e;

    double s3 = c3 * dsm::log(r3);
    double s3_h = s3 * c3_h / c3 + c3 * r3_h / r3;
    double s3_hh = (s3_h * c3_h + s3 * c3_hh) / c3 - s3 * ipow(c3_h / c3, 2) +
                   (c3_h * r3_h + c3 * r3_hh) / r3 - c3 * ipow(r3_h / r3, 2);
    double s3_g = c3 * r3_g / r3;
    double s3_gg = c3 * (r3_gg / r3 - ipow(r3_g / r3, 2));
    double s3_hg = (c3_h * r3_g + c3 * r3_hg) / r3 - c3 * r3_g * r3_h / (r3 * r3);

    double b4 = 2.0 - d * d / e;
    double b4_h = e_h * ipow(d / e, 2) - 2.0 * d * d_h / e;
    double b4_hh = e_hh * ipow(d / e, 2) + 2.0 * e_h * ipow(d / e, 2) * (d_h / d - e_h / e) -
                   2.0 * (d_h * d_h + d * d_hh) / e + 2. * d * d_h * e_h / (e * e);

    double c4 = 2.0 * e * sqge + d;
    double c4_h = 2.0 * e_h * sqge + d_h;
    double c4_hh = 2.0 * e_hh * sqge + d_hh;
    double c4_g = e / sqge;
    double c4_gg = -0.5 * e / (Gamma_e * sqge);
    double c4_hg = e_h / sqge;

    double s4a = 2.0 / e / discr;
    double s4ah = e_h / e + di_h / discr;
    double s4a_h = -s4a * s4ah;
    double s4a_hh = -s4a_h * s4ah - s4a * (e_hh / e - ipow(e_h / e, 2) + di_hh / discr - ipow(di_h / discr, 2));
    double s4b = d * b3 + b4 * b2;
    double s4b_h = d_h * b3 + d * b3_h + b4_h * b2 + b4 * b2_h;
    double s4b_hh = d_hh * b3 + 2.0 * d_h * b3_h + d * b3_hh + b4_hh * b2 + 2. * b4_h * b2_h + b4 * b2_hh;
    double s4c = dsm::atan(c4 / discr) - dsm::atan(d / discr);
    double up1 = c4_h * discr - c4 * di_h;
    double dn1 = discr * discr + c4 * c4;
    double up2 = d_h * discr - d * di_h;
    double dn2 = discr * discr + d * d;
    double s4c_h = up1 / dn1 - up2 / dn2;
    double s4c_hh = (c4_hh * discr - c4 * di_hh) / dn1 - up1 * 2.0 * (discr * di_h + c4 * c4_h) / (dn1 * dn1) -
                    (d_hh * discr - d * di_hh) / dn2 + up2 * 2.0 * (discr * di_h + d * d_h) / (dn2 * dn2);
    double s4c_g = c4_g * discr / dn1;
    double s4c_gg = c4_gg * discr / dn1 - 2.0 * c4 * discr * ipow(c4_g / dn1, 2);
    double s4c_hg = (c4_hg * discr + c4_g * di_h - c4_g * discr / dn1 * 2. * (discr * di_h + c4 * c4_h)) / dn1;
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
    if (std::fabs(s) < R4(1.0e-9) * std::fabs(theta * f_h)) s = 0.0;
    double cv = 2.0 * theta * (Gamma_e * f_hg - f_h) - theta * theta * f_hh - Gamma_e * Gamma_e * f_gg;
    if (std::fabs(cv) < R4(1.0e-9) * std::fabs(Gamma_e * Gamma_e * f_gg)) cv = 0.0;
    double pdlh = onethird * theta * (Gamma_e * f_hg - 2.0 * f_h - 2.0 * theta * f_hh);
    double pdlg = onethird * Gamma_e * (f_g + Gamma_e * f_gg - 2.0 * theta * f_hg);
    double dpr = p + onethird * (pdlg - 2.0 * pdlh);
    double dpt = Gamma_e * (theta * f_hg - onethird * Gamma_e * f_gg) - theta * (f_h / 0.75 + theta * f_hh / 1.5);
    r = Thermo{f, u, p, s, cv, dpr, dpt};
}

// ------------------------------------------------------------------ ions
namespace {

// ion.f:178-285
void ie_liquid(double Gamma_e, double rs, double Z, Thermo& r) {
    static const double s3 = (double)std::sqrt(3.0f);
    static const double CTFfac = (double)(18.0f / 175.0f) * dsm::pow(12.0 / pi, 2.0 * onethird);
    double Z13 = dsm::pow(Z, onethird);
    double cDH = (Z / s3) * (dsm::pow(1.0 + Z, 1.5) - 1.0 - dsm::pow(Z, 1.5));
    double cTF = CTFfac * (Z * Z) * (Z13 - 1.0 + R4(0.2) / std::sqrt(Z13));
    double xrs = finestructure * dsm::pow(9.0 * pi / 4.0, onethird);
    double xr = xrs / rs;
    double lZ = dsm::log(Z);

    double a = R4(1.11) * dsm::pow(Z, R4(0.475));
    double b = R4(0.2) + R4(0.078) * (lZ * lZ);
    double nu = R4(1.16) + R4(0.08) * lZ;
    double sGam = std::sqrt(Gamma_e);
    double nuGam = dsm::pow(Gamma_e, nu);
    double nuGam_g = nu * nuGam / Gamma_e;
    double nuGam_gg = (nu - 1.0) * nuGam_g / Gamma_e;
    double ty1 = 1.0 / (R4(0.001) * (Z * Z) + 2.0 * Gamma_e);
    double ty1_g = -2.0 * (ty1 * ty1);
    double ty1_gg = -4.0 * ty1 * ty1_g;
    double rs2 = rs * rs, rs3 = rs * rs * rs;
    double ty2 = 1.0 + 6.0 * rs2;
    double ty2_x = -12.0 * rs2 / xr;
    double ty2_xx = -3.0 * ty2_x / xr;
    double ty = rs3 / ty2 * (1.0 + ty1);
    double tyx = 3.0 / xr + ty2_x / ty2;
    double ty_x = -ty * tyx;
    double ty_g = rs3 * ty1_g / ty2;
    double p1 = (Z - 1.0) / 9.0;
    double cor1 = 1.0 + p1 * ty;
    double cor1_x = p1 * ty_x;
    double cor1_g = p1 * ty_g;
    double cor1_xx = p1 * (ty * (3.0 / (xr * xr) + ipow(ty2_x / ty2, 2) - ty2_xx / ty2) - ty_x * tyx);
    double cor1_gg = p1 * rs3 * ty1_gg / ty2;
    double cor1_xg = -p1 * ty_g * tyx;
    double u0 = R4(0.78) * std::sqrt(Gamma_e / Z) * rs3;
    double u0_x = -3.0 * u0 / xr;
    double u0_g = 0.5 * u0 / Gamma_e;
    double u0_xx = -4.0 * u0_x / xr;
    double u0_gg = -0.5 * u0_g / Gamma_e;
    double u0_xg = -3.0 * u0_g / xr;
    double d0_g = Z * Z * Z;
    double d0 = Gamma_e * d0_g + 21.0 * rs3;
    double d0_x = -63.0 * rs3 / xr;
    double d0_xx = 252.0 * rs3 / (xr * xr);
    double cor0 = 1.0 + u0 / d0;
    double cor0_x = (u0_x - u0 * d0_x / d0) / d0;
    double cor0_g = (u0_g - u0 * d0_g / d0) / d0;
    double cor0_xx = (u0_xx - (2.0 * u0_x * d0_x + u0 * d0_xx) / d0 + 2.0 * ipow(d0_x / d0, 2)) / d0;
    double cor0_gg = (u0_gg - 2.0 * u0_g * d0_g / d0 + 2.0 * u0 * ipow(d0_g / d0, 2)) / d0;
    double cor0_xg = (u0_xg - (u0_x * d0_g + u0_g * d0_x) / d0 + 2.0 * u0 * d0_x * d0_g / (d0 * d0)) / d0;

    double rgam = std::sqrt(1.0 + xr * xr);
    double q1 = R4(0.18) / dsm::pow(Z, 0.25);
    double q2 = R4(0.2) + R4(0.37) / std::sqrt(Z);
    double h1u = 1.0 + R4(0.2) * (xr * xr);
    double h1d = 1.0 + q1 * xr + q2 * (xr * xr);
    double h1 = h1u / h1d;
    double h1x = R4(0.4) * xr / h1u - (q1 + 2. * q2 * xr) / h1d;
    double h1_x = h1 * h1x;
    double h1_xx = h1_x * h1x + h1 * (R4(.4) / h1u - ipow(R4(.4) * xr / h1u, 2) - 2. * q2 / h1d +
                                      ipow((q1 + 2. * q2 * xr) / h1d, 2));
    double up = cDH * sGam + a * cTF * nuGam * cor0 * h1;
    double up_x = a * cTF * nuGam * (cor0_x * h1 + cor0 * h1_x);
    double up_g = 0.5 * cDH / sGam + a * cTF * (nuGam_g * cor0 + nuGam * cor0_g) * h1;
    double up_xx = a * cTF * nuGam * (cor0_xx * h1 + 2. * cor0_x * h1_x + cor0 * h1_xx);
    double up_gg = -0.25 * cDH / (sGam * Gamma_e) +
                   a * cTF * (nuGam_gg * cor0 + 2.0 * nuGam_g * cor0_g + nuGam * cor0_gg) * h1;
    double up_xg = a * cTF * (nuGam_g * (cor0_x * h1 + cor0 * h1_x) + nuGam * (cor0_xg * h1 + cor0_g * h1_x));
    double dn1 = b * sGam + a / rs * nuGam * cor1;
    double dn1_x = a * nuGam * (cor1 / xrs + cor1_x / rs);
    double dn1_g = 0.5 * b / sGam + a / rs * (nuGam_g * cor1 + nuGam * cor1_g);
    double dn1_xx = a * nuGam / xrs * (2. * cor1_x + xr * cor1_xx);
    double dn1_gg = -0.25 * b / (Gamma_e * sGam) +
                    a / rs * (nuGam_gg * cor1 + 2. * nuGam_g * cor1_g + nuGam * cor1_gg);
    double dn1_xg = a * (nuGam_g * (cor1 / xrs + cor1_x / rs) + nuGam * (cor1_g / xrs + cor1_xg / rs));
    double dn = 1.0 + dn1 / rgam;
    double dn_x = dn1_x / rgam - xr * dn1 / ipow(rgam, 3);
    double dn_xx = (dn1_xx - ((2. * xr * dn1_x + dn1) - 3. * (xr * xr) * dn1 / (rgam * rgam)) / (rgam * rgam)) / rgam;
    double dn_g = dn1_g / rgam;
    double dn_gg = dn1_gg / rgam;
    double dn_xg = dn1_xg / rgam - xr * dn1_g / ipow(rgam, 3);
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
    double dpr = (12.0 * p + (xr * xr) * f_xx + 2.0 * xr * Gamma_e * f_xg + (Gamma_e * Gamma_e) * f_gg) / 9.0;
    r = Thermo{f, u, p, s, cv, dpr, dpt};
}

// ion.f:287-405
void ie_lattice(double Gamma, double theta, double rs, double Z, Thermo& r) {
    const double a[4] = {R4(1.1866), R4(0.684), R4(17.9), R4(41.5)};
    const double px = R4(0.205);
    double e1 = dsm::exp(1.0);
    double aTF = (double)(54.0f / 175.0f) * dsm::pow(12.0 / pi, onethird) * finestructure;
    double xrs = finestructure * dsm::pow(9.0 * pi / 4.0, onethird);
    double xr = xrs / rs;
    double lZ = dsm::log(Z);
    double r3 = 1.0 / (1.0 + R4(0.01) * dsm::pow(lZ, 1.5) + R4(0.097) / (Z * Z));
    double b[4];
    b[0] = 1.0 - a[0] / dsm::pow(Z, R4(0.267)) + R4(0.27) / Z;
    b[1] = 1.0 + 2.25 / dsm::pow(Z, onethird) * (1.0 + a[1] * ipow(Z, 5) + R4(0.222) * ipow(Z, 6)) /
                     (1.0 + R4(0.222) * ipow(Z, 6));
    b[2] = a[3] / (1.0 + lZ);
    b[3] = R4(0.395) * lZ + R4(0.347) / dsm::pow(Z, 1.5);

    double finf = aTF * dsm::pow(Z, twothird) * b[0] * std::sqrt(1.0 + b[1] / (xr * xr));
    double finfx = -b[1] / ((b[1] + xr * xr) * xr);
    double finf_x = finf * finfx;
    double finf_xx = finf_x * finfx - finf_x * (b[1] + 3.0 * (xr * xr)) / ((b[1] + xr * xr) * xr);

    double q1u = b[2] + a[2] * (xr * xr);
    double q1d = 1.0 + b[3] * (xr * xr);
    double q1 = q1u / q1d;
    double q1x = 2.0 * xr * (a[2] / q1u - b[3] / q1d);
    double q1x_x = q1x / xr + 4.0 * (xr * xr) * (ipow(b[3] / q1d, 2) - ipow(a[2] / q1u, 2));
    double q1_x = q1 * q1x;
    double q1_xx = q1_x * q1x + q1 * q1x_x;

    double sup, sup_x, sup_g, sup_xx, sup_gg, sup_xg;
    if (theta < 6.0 / px) {
        double y0 = ipow(px * theta, 2);
        double y0_x = y0 / xr;
        double y0_g = 2.0 * y0 / Gamma;
        double y0_xx = 0.0;
        double y0_gg = y0_g / Gamma;
        double y0_xg = y0_g / xr;
        double y1 = dsm::exp(y0);
        double y1_x = y1 * y0_x;
        double y1_g = y1 * y0_g;
        double y1_xx = y1 * (y0_x * y0_x + y0_xx);
        double y1_gg = y1 * (y0_g * y0_g + y0_gg);
        double y1_xg = y1 * (y0_x * y0_g + y0_xg);
        double sa = 1.0 + y1;
        double supa = dsm::log(sa);
        double supa_x = y1_x / sa;
        double supa_g = y1_g / sa;
        double supa_xx = (y1_xx - y1_x * y1_x / sa) / sa;
        double supa_gg = (y1_gg - y1_g * y1_g / sa) / sa;
        double supa_xg = (y1_xg - y1_x * y1_g / sa) / sa;
        double em2 = e1 - 2.0;
        double sb = e1 - em2 / y1;
        double supb = dsm::log(sb);
        double em2y1 = em2 / ((y1 * y1) * sb);
        double supb_x = em2y1 * y1_x;
        double supb_g = em2y1 * y1_g;
        double supb_xx = em2y1 * (y1_xx - 2.0 * (y1_x * y1_x) / y1 - y1_x * supb_x);
        double supb_gg = em2y1 * (y1_gg - 2.0 * (y1_g * y1_g) / y1 - y1_g * supb_g);
        double supb_xg = em2y1 * (y1_xg - 2.0 * y1_x * y1_g / y1 - y1_g * supb_x);
        sup = std::sqrt(supa / supb);
        double supx = 0.5 * (supa_x / supa - supb_x / supb);
        sup_x = sup * supx;
        double supg = 0.5 * (supa_g / supa - supb_g / supb);
        sup_g = sup * supg;
        sup_xx = sup_x * supx + sup * 0.5 * (supa_xx / supa - ipow(supa_x / supa, 2) - supb_xx / supb +
                                             ipow(supb_x / supb, 2));
        sup_gg = sup_g * supg + sup * 0.5 * (supa_gg / supa - ipow(supa_g / supa, 2) - supb_gg / supb +
                                             ipow(supb_g / supb, 2));
        sup_xg = sup_x * supg + sup * 0.5 * ((supa_xg - supa_x * supa_g / supa) / supa -
                                             (supb_xg - supb_x * supb_g / supb) / supb);
    } else {
        sup = px * theta;
        sup_x = 0.5 * px * theta / xr;
        sup_g = px * theta / Gamma;
        sup_xx = -0.5 * sup_x / xr;
        sup_gg = 0.0;
        sup_xg = sup_x / Gamma;
    }
    double gr3 = dsm::pow(Gamma / sup, r3);
    double gr3x = -r3 * sup_x / sup;
    double gr3_x = gr3 * gr3x;
    double gr3_xx = gr3_x * gr3x - r3 * gr3 * (sup_xx / sup - ipow(sup_x / sup, 2));
    double gr3g = r3 * (1.0 / Gamma - sup_g / sup);
    double gr3_g = gr3 * gr3g;
    double gr3_gg = gr3_g * gr3g + gr3 * r3 * (ipow(sup_g / sup, 2) - sup_gg / sup - 1.0 / (Gamma * Gamma));
    double gr3_xg = gr3_g * gr3x + gr3 * r3 * (sup_x * sup_g / (sup * sup) - sup_xg / sup);
    double w = 1.0 + q1 / gr3;
    double w_x = q1_x / gr3 - q1 * gr3_x / (gr3 * gr3);
    double w_g = -q1 * gr3_g / (gr3 * gr3);
    double w_xx = q1_xx / gr3 - (2.0 * q1_x * gr3_x + q1 * (gr3_xx - 2.0 * (gr3_x * gr3_x) / gr3)) / (gr3 * gr3);
    double w_gg = q1 * (2.0 * (gr3_g * gr3_g) / gr3 - gr3_gg) / (gr3 * gr3);
    double w_xg = -(q1_x * gr3_g + q1 * (gr3_xg - 2.0 * gr3_x * gr3_g / gr3)) / (gr3 * gr3);

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
    double dpr = (12.0 * p + (xr * xr) * f_xx + 2.0 * xr * Gamma * f_xg + (Gamma * Gamma) * f_gg) / 9.0;
    r = Thermo{f, u, p, s, cv, dpr, dpt};
}

// ion.f:407-452
void ii_liquid(double Gamma, double theta, Thermo& r) {
    // array constructor of REAL(4): A(3) is evaluated entirely in single precision
    static const float A1f = -0.907347f, A2f = 0.62849f;
    static const float A3f = -0.5f * std::sqrt(3.0f) + 0.907347f / std::sqrt(0.62849f);
    const double A1 = (double)A1f, A2 = (double)A2f, A3 = (double)A3f;
    const double B1 = R4(4.50e-3), B2 = R4(170.0), B3 = R4(-8.4e-5), B4 = R4(3.70e-3);

    double f0 = A1 * (std::sqrt(Gamma * (A2 + Gamma)) -
                      A2 * dsm::log(std::sqrt(Gamma / A2) + std::sqrt(1.0 + Gamma / A2))) +
                2.0 * A3 * (std::sqrt(Gamma) - dsm::atan(std::sqrt(Gamma)));
    double G2 = Gamma * Gamma;
    double G15 = dsm::pow(Gamma, 1.5);
    double f = f0 + B1 * (Gamma - B2 * dsm::log(1.0 + Gamma / B2)) + 0.5 * B3 * dsm::log(1.0 + G2 / B4);
    double u = G15 * (A1 / std::sqrt(A2 + Gamma) + A3 / (1.0 + Gamma)) +
               G2 * (B1 / (B2 + Gamma) + B3 / (B4 + G2));
    double cv = 0.5 * G15 * (A3 * (Gamma - 1.0) / ipow(Gamma + 1.0, 2) - A1 * A2 / dsm::pow(Gamma + A2, 1.5)) +
                G2 * (B3 * (G2 - B4) / ipow(G2 + B4, 2) - B1 * B2 / ipow(Gamma + B2, 2));
    double p = onethird * u;
    double dpr = (4.0 * u - cv) / 9.0;
    double dpt = onethird * cv;

    double fid = 3.0 * dsm::log(theta) - 1.5 * dsm::log(Gamma) - 0.5 * dsm::log(6.0 / pi) - 1.0;
    f = f + fid;
    u = u + 1.5;
    p = p + 1.0;
    cv = cv + 1.5;
    double s = u - f;
    dpt = dpt + 1.0;
    dpr = dpr + 1.0;

    double fq = theta * theta / 24.0;
    f = f + fq;
    u = u + 2.0 * fq;
    s = s + fq;
    cv = cv - 2.0 * fq;
    p = p + fq;
    dpr = dpr + 2.0 * fq;
    dpt = dpt - fq;
    r = Thermo{f, u, p, s, cv, dpr, dpt};
}

// ion.f:454-576 ; returns the `total` component
void ii_lattice(double Gamma, double theta, Thermo& tot) {
    const double KM = R4(-0.895929255682), w = R4(0.5113875), clm = R4(-2.49389);
    const double alpha[9] = {R4(0.0), R4(0.932446), R4(0.334547), R4(0.265764), R4(0.0),
                             R4(0.0), R4(4.757014e-3), R4(0.0), R4(4.7770935e-3)};
    const double a[9] = {R4(1.0), R4(0.1839), R4(0.593586), R4(5.4814e-3), R4(5.01813e-4),
                         R4(0.0), R4(3.9247e-7), R4(0.0), R4(5.8356e-11)};
    const double b[8] = {R4(261.66), R4(0.0), R4(7.07997), R4(0.0), R4(0.0409484),
                         R4(3.97355e-4), R4(5.11148e-5), R4(2.19749e-6)};
    const double afh[4] = {0.0, R4(10.9), R4(247.0), R4(1.765e5)};  // 1-based
    const double b1 = R4(0.12);
    const double c1 = b1 / afh[1];
    const double theta_low = R4(1.0e-5), theta_high = R4(1.0e5);

    double fh = 0, uh = 0, cvh = 0;
    if (theta > theta_high) {
        uh = 3.0 / (alpha[6] * ipow(theta, 3));
        fh = -uh / 3.0;
        cvh = 4.0 * uh;
    } else if (theta < theta_low) {
        fh = 3.0 * dsm::log(theta) + clm - 1.5 * w * theta + theta * theta / 24.0;
        uh = 3.0 - 1.5 * w * theta + theta * theta / 12.0;
        cvh = 3.0 - theta * theta / 12.0;
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
        Bc[0] = Bc[0] + alpha[6] * a[6] * ipow(theta, 9) + alpha[8] * a[8] * ipow(theta, 11);
        Bc[1] = Bc[1] + 9.0 * alpha[6] * a[6] * ipow(theta, 8) + 11.0 * alpha[8] * a[8] * ipow(theta, 10);
        Bc[2] = Bc[2] + 72.0 * alpha[6] * a[6] * ipow(theta, 7) + 110.0 * alpha[8] * a[8] * ipow(theta, 9);
        for (int n = 1; n <= 3; ++n) {
            double ea = dsm::exp(-alpha[n] * theta);
            double omea = 1.0 - ea;
            fh = fh + dsm::log(omea);
            uh = uh + alpha[n] * ea / omea;
            cvh = cvh + alpha[n] * alpha[n] * ea / (omea * omea);
        }
        fh = fh - Ac[0] / Bc[0];
        uh = (uh - (Ac[1] * Bc[0] - Ac[0] * Bc[1]) / (Bc[0] * Bc[0])) * theta;
        cvh = cvh + ((Ac[2] * Bc[0] - Bc[2] * Ac[0]) * Bc[0] - 2.0 * (Ac[1] * Bc[0] - Bc[1] * Ac[0]) * Bc[1]) /
                        ipow(Bc[0], 3);
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
    double theta4 = ipow(theta, 4);
    double c1t2 = c1 * theta2;
    double sup = dsm::exp(-c1t2);
    double tq = b1 * theta2 / Gamma;
    double gninv = sup;
    for (int n = 1; n <= 3; ++n) {
        gninv = gninv / Gamma;
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
    dpra = dpra - tq / 4.5;

    double madelung = KM * Gamma;
    tot.f = fh + fa + madelung;
    tot.u = uh + ua + madelung;
    tot.p = ph + pa + onethird * madelung;
    tot.s = sh;
    tot.cv = cvh + cva;
    tot.dpt = dpth + dpta;
    tot.dpr = dprh + dpra + madelung / 2.25;
}

// ion.f:135-176
void one_ion(double Gamma_e, double rs, double Z, double A, int phase, Thermo& r) {
    static const double sevensixth = (double)(7.0f / 6.0f);
    double Gamma = Gamma_e * dsm::pow(Z, fivethird);
    double theta = Gamma * std::sqrt(3.0 * Melectron / A / amu / rs) / dsm::pow(Z, sevensixth);
    Thermo ie, ii;
    if (phase == liquid_phase) {
        ie_liquid(Gamma_e, rs, Z, ie);
        ii_liquid(Gamma, theta, ii);
    } else {
        ie_lattice(Gamma, theta, rs, Z, ie);
        ii_lattice(Gamma, theta, ii);
    }
    r.f = ie.f + ii.f; r.u = ie.u + ii.u; r.p = ie.p + ii.p; r.s = ie.s + ii.s;
    r.cv = ie.cv + ii.cv; r.dpr = ie.dpr + ii.dpr; r.dpt = ie.dpt + ii.dpt;
}

// ion.f:94-133
void LM_corrections(double rs, double Gamma_e, const Composition& c, Thermo& r) {
    const double threshold = R4(1.0e-9);
    double Gamma = Gamma_e * c.Z53;
    double dif0;
    if (rs < threshold) dif0 = c.Z52 - std::sqrt(ipow(c.Z2, 3) / c.Z);
    else dif0 = c.ZZ1_32 - std::sqrt(ipow(c.Z2 + c.Z, 3) / c.Z);
    double difR = dif0 / c.Z52;
    double difFDH = dif0 * Gamma_e * std::sqrt(onethird * Gamma_e);
    double D = c.Z2 / (c.Z * c.Z);
    r = Thermo{0, 0, 0, 0, 0, 0, 0};
    if (std::fabs(D - 1.0) < threshold) return;
    double p3 = dsm::pow(D, R4(-0.2));
    double d0 = (R4(2.6) * difR + 14.0 * ipow(difR, 3)) / (1.0 - p3);
    double gp = d0 * dsm::pow(Gamma, p3);
    double f0 = difFDH / (1.0 + gp);
    double q = (D * D) * R4(0.0117);
    double rr = 1.5 / p3 - 1.0;
    double gq = q * gp;
    double f = f0 / dsm::pow(1.0 + gq, rr);
    double g = 1.5 - p3 * gp / (1.0 + gp) - rr * p3 * gq / (1.0 + gq);
    double u = f * g;
    double s = u - f;
    double p = onethird * u;
    double g_dg = -(p3 * p3) * (gp / ipow(1.0 + gp, 2) + rr * gq / ipow(1.0 + gq, 2));
    double u_dg = u * g + f * g_dg;
    double cv = u - u_dg;
    double dpr = p + u_dg / 9.0;
    double dpt = p - onethird * u_dg;
    r = Thermo{f, u, p, s, cv, dpr, dpt};
}
}  // namespace

// ion.f:20-92 ; return value is `err` (0 or strong_quantum_effects = 1)
int ion_mixture(const EosControls& rq, double rs, double Gamma_e, const Composition& c,
                int ncharged, const double* Zion, const double* Aion, const double* Yion,
                double& Gamma, double& ionQ, int& phase, Thermo& r) {
    const double Q2_threshold = 18.0;
    int err = 0;
    double Mu = amu / Melectron;
    Gamma = Gamma_e * c.Z53;
    phase = (Gamma < rq.Gamma_melt) ? liquid_phase : solid_phase;
    double ionQ2 = 3.0 * (Gamma_e * Gamma_e) / rs / Mu * c.Z2XoA2 * c.A / c.Z;
    ionQ = std::sqrt(ionQ2);
    if (phase == liquid_phase && ionQ2 > Q2_threshold) err = 1;

    double f = 0, u = 0, p = 0, s = 0, cv = 0, dpr = 0, dpt = 0;
    for (int i = 0; i < ncharged; ++i) {
        if (Yion[i] > rq.Ythresh) {
            Thermo t;
            one_ion(Gamma_e, rs, Zion[i], Aion[i], phase, t);
            f = f + Yion[i] * t.f;
            u = u + Yion[i] * t.u;
            s = s + Yion[i] * (t.s - dsm::log(Yion[i]));
            p = p + Yion[i] * t.p;
            cv = cv + Yion[i] * t.cv;
            dpr = dpr + Yion[i] * t.dpr;
            dpt = dpt + Yion[i] * t.dpt;
        }
    }
    Thermo m;
    LM_corrections(rs, Gamma_e, c, m);
    r.f = f * c.A + m.f;
    r.u = u * c.A + m.u;
    r.s = s * c.A + m.s;
    r.p = p * c.A + m.p;
    r.cv = cv * c.A + m.cv;
    r.dpr = dpr * c.A + m.dpr;
    r.dpt = dpt * c.A + m.dpt;
    return err;
}

// ------------------------------------------------------- Fermi-Dirac fits
#include "fermi_coef.inc"
namespace {
inline double ratfit(const double* a, int m, const double* b, int k, double xx) {
    // rn = xx + a(m); rn = rn*xx + a(i), i=m-1..1 ; den = b(k+1); den = den*xx + b(i), i=k..1
    double rn = xx + a[m - 1];
    for (int i = m - 2; i >= 0; --i) rn = rn * xx + a[i];
    double den = b[k];
    for (int i = k - 1; i >= 0; --i) den = den * xx + b[i];
    return rn / den;
}
}  // namespace
double zfermim12(double x) {  // funct_fermi1.f:21-86
    if (x < 2.0) { double xx = dsm::exp(x); return xx * ratfit(zfermim12_a1, 7, zfermim12_b1, 7, xx); }
    double xx = 1.0 / (x * x);
    return std::sqrt(x) * ratfit(zfermim12_a2, 11, zfermim12_b2, 11, xx);
}
double zfermi12(double x) {  // :88-146
    if (x < 2.0) { double xx = dsm::exp(x); return xx * ratfit(zfermi12_a1, 7, zfermi12_b1, 7, xx); }
    double xx = 1.0 / (x * x);
    return x * std::sqrt(x) * ratfit(zfermi12_a2, 10, zfermi12_b2, 11, xx);
}
double zfermi32(double x) {  // :216-276
    if (x < 2.0) { double xx = dsm::exp(x); return xx * ratfit(zfermi32_a1, 6, zfermi32_b1, 7, xx); }
    double xx = 1.0 / (x * x);
    return x * x * std::sqrt(x) * ratfit(zfermi32_a2, 9, zfermi32_b2, 10, xx);
}
double ifermi12(double f) {  // :528-580
    const double an = 0.5;
    if (f < 4.0) return dsm::log(f * ratfit(ifermi12_a1, 4, ifermi12_b1, 3, f));
    double ff = 1.0 / dsm::pow(f, 1.0 / (1.0 + an));
    double rn = ff + ifermi12_a2[5];
    for (int i = 4; i >= 0; --i) rn = rn * ff + ifermi12_a2[i];
    double den = ifermi12_b2[5];
    for (int i = 4; i >= 0; --i) den = den * ff + ifermi12_b2[i];
    return rn / (den * ff);
}

// neutron.f:15-69
void MB77(double n, double T, double Tns, Thermo& r) {
    const double c[4] = {R4(1.2974), R4(15.0298), R4(-15.2343), R4(7.4663)};
    r = Thermo{0, 0, 0, 0, 0, 0, 0};
    if (n == 0.0) return;
    double k = dsm::pow(0.5 * threepisquare * n, onethird) / cm_to_fm;
    double dkdn = onethird * k / n;
    double u = k * (c[0] + k * (c[1] + k * (c[2] + k * c[3])));
    u = u * mev_to_ergs;
    double p = onethird * n * k * (c[0] + k * (2.0 * c[1] + k * (3.0 * c[2] + k * 4.0 * c[3])));
    p = p * mev_to_ergs;
    double dpdk = onethird * n * k * (c[0] + k * (4.0 * c[1] + k * (9.0 * c[2] + 16.0 * c[3] * k))) * mev_to_ergs;
    double dpr = (p / n + dpdk * dkdn) * n;
    double dpT = 0.0;

    double lambda3 = pi * pi * dsm::pow(hbar * hbar / (Mneutron * boltzmann * T), 1.5) / (double)std::sqrt(2.0f);
    double zeta = ifermi12(n * lambda3);
    double cv = 2.5 * zfermi32(zeta) / zfermi12(zeta) - 4.5 * zfermi12(zeta) / zfermim12(zeta);
    cv = cv * boltzmann;
    double s = fivethird * zfermi32(zeta) / zfermi12(zeta) - zeta;
    s = s * boltzmann;
    double R;
    if (T < Tns) {
        double tau = T / Tns;
        double v = std::sqrt(1.0 - tau) * (R4(1.456) - R4(0.157) / std::sqrt(tau) + R4(1.764) / tau);
        R = dsm::pow(R4(0.4186) + std::sqrt((double)(1.007f * 1.007f) + ipow(R4(0.5010) * v, 2)), 2.5) *
            dsm::exp(R4(1.456) - std::sqrt((double)(1.456f * 1.456f) + v * v));
    } else {
        R = 1.0;
    }
    cv = cv * R;
    s = s * R;
    double f = u - T * s;
    r = Thermo{f, u, p, s, cv, dpr, dpT};
}

// radiation.f:5-25
void get_radiation_eos(double rho, double T, Thermo& r) {
    const double fourthird = 4.0 * onethird;
    double aT3 = arad * ipow(T, 3);
    double aT4 = arad * ipow(T, 4);
    r.u = aT4 / rho;
    r.p = onethird * aT4;
    r.f = -onethird * aT4 / rho;
    r.s = fourthird * aT3 / rho;
    r.dpr = 0.0;
    r.dpt = fourthird * aT4;
    r.cv = 4.0 * aT3 / rho;
}

// dStar_eos_lib.f:123-297
int eval_crust_eos(const HelmTable& h, const EosControls& rq, double rho, double T,
                   const Composition& c, int ncharged, const double* Zion, const double* Aion,
                   const double* Yion, const double Tcs[3], double res[num_eos_results],
                   int& phase, double& chi, Thermo comps[4]) {
    double ne, rs, Gamma_e, Gamma, ionQ;
    lattice_properties(rho, T, c, ne, rs, Gamma_e, Gamma, ionQ);
    if (chi == use_default_nuclear_size) chi = nuclear_volume_fraction(rho, c, default_nuclear_radius);

    Thermo e, ex, ion, nt, rad;
    double eta_e = 0.0;
    int ierr_h = get_helm_eos_results(h, rho, T, c, e, eta_e);
    ee_exchange(Gamma_e, rs, ex);
    double nek = ne * boltzmann;
    double nekT = nek * T;
    double uexfac = nekT / rho, pexfac = nekT, sexfac = nek / rho;

    int ierr = ion_mixture(rq, rs, Gamma_e, c, ncharged, Zion, Aion, Yion, Gamma, ionQ, phase, ion);
    double n_i = ne / c.Z;
    double nik = n_i * boltzmann;
    double nikT = nik * T;
    double uifac = nikT / rho, pifac = nikT, sifac = nik / rho;
    (void)ierr;

    if (!(chi < 1.0)) return -10;  // neutron_volume assertion
    double nn = rho * c.Yn / amu / (1.0 - chi);
    double Tns = Tcs[1];
    MB77(nn, T, Tns, nt);
    double unfac = avogadro * c.Yn;
    double pnfac = 1.0;
    double snfac = unfac;

    get_radiation_eos(rho, T, rad);

    double p = e.p + ex.p * pexfac + ion.p * pifac + nt.p * pnfac + rad.p;
    double u = e.u + ex.u * uexfac + ion.u * uifac + nt.u * unfac + rad.u;
    double s = e.s + ex.s * sexfac + ion.s * sifac + nt.s * snfac + rad.s;
    double cv = e.cv + ex.cv * sexfac + ion.cv * sifac + nt.cv * snfac + rad.cv;
    double dpr = e.dpr + ex.dpr * pexfac + ion.dpr * pifac + nt.dpr * pnfac + rad.dpr;
    double dpt = e.dpt + ex.dpt * pexfac + ion.dpt * pifac + nt.dpt * pnfac + rad.dpt;
    double gamma3m1 = dpt / rho / T / cv;
    double gamma1 = (dpr + dpt * gamma3m1) / p;
    double grad_ad = gamma3m1 / gamma1;
    double cp = gamma1 * p * cv / dpr;
    double mu_e = (eta_e * boltzmann * T) * ergs_to_mev;
    double mu_n;
    if (c.Yn < rq.Ythresh) mu_n = 0.0;
    else mu_n = (nt.u - T * nt.s + nt.p / nn) * ergs_to_mev;

    res[i_lnP] = dsm::log(p);
    res[i_lnE] = dsm::log(u);
    res[i_lnS] = dsm::log(s);
    res[i_grad_ad] = grad_ad;
    res[i_chiRho] = dpr / p;
    res[i_chiT] = dpt / p;
    res[i_Cp] = cp;
    res[i_Cv] = cv;
    res[i_Chi] = chi;
    res[i_Gamma] = Gamma;
    res[i_Theta] = 1.0 / ionQ;
    res[i_mu_e] = mu_e;
    res[i_mu_n] = mu_n;
    res[i_Gamma1] = gamma1;
    res[i_Gamma3] = gamma3m1;
    if (comps) {
        comps[0] = Thermo{e.f + ex.f * uexfac, e.u + ex.u * uexfac, e.p + ex.p * pexfac,
                          e.s + ex.s * sexfac, e.cv + ex.cv * sexfac,
                          (e.dpr + ex.dpr * pexfac) / (e.p + ex.p * pexfac),
                          (e.dpt + ex.dpt * pexfac) / (e.p + ex.p * pexfac)};
        comps[1] = Thermo{ion.f * uifac, ion.u * uifac, ion.p * pifac, ion.s * sifac, ion.cv * sifac,
                          ion.dpr / ion.p, ion.dpt / ion.p};
        if (c.Yn < rq.Ythresh) comps[2] = Thermo{0, 0, 0, 0, 0, 0, 0};
        else comps[2] = Thermo{nt.f * unfac, nt.u * unfac, nt.p * pnfac, nt.s * snfac, nt.cv * snfac,
                               nt.dpr / nt.p, nt.dpt / nt.p};
        comps[3] = Thermo{rad.f, rad.u, rad.p, rad.s, rad.cv, rad.dpr / rad.p, rad.dpt / rad.p};
    }
    return ierr_h;
}

}  // namespace dso
