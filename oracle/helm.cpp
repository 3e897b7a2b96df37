// ORACLE (test infrastructure): tabular Helmholtz free-energy electron-positron EOS.
// Restates dStar_eos/private/helm_alloc.f:120-330 (table geometry), helm_core.f:12-158
// (helmeos2 blending logic, only the alfa=0 branch logT>=5 is reachable from NScool),
// helm_core.f:161-503 and helm_electron_positron.dek:1-486 (biquintic/bicubic Hermite).
#include <algorithm>
#include <cmath>

#include "dstar_oracle.hpp"

namespace dso {

void helm_setup(HelmTable& h, int imax, int jmax, const double* const a[21]) {
    h.imax = imax;
    h.jmax = jmax;
    h.logtlo = 3.0; h.logthi = 13.0; h.logdlo = -12.0; h.logdhi = 14.0;  // helm_alloc.f:150-153
    h.templo = 1.0e3;
    h.temphi = 1.0e13;
    h.denlo = 1.0e-12;
    h.denhi = 1.0e14;
    h.logtstp = (h.logthi - h.logtlo) / (double)(float)(jmax - 1);
    h.logtstpi = 1.0 / h.logtstp;
    h.logdstp = (h.logdhi - h.logdlo) / (double)(float)(imax - 1);
    h.logdstpi = 1.0 / h.logdstp;
    size_t n = (size_t)imax * jmax;
    std::vector<double>* dst[21] = {&h.f, &h.fd, &h.ft, &h.fdd, &h.ftt, &h.fdt, &h.fddt,
                                    &h.fdtt, &h.fddtt, &h.dpdf, &h.dpdfd, &h.dpdft, &h.dpdfdt,
                                    &h.ef, &h.efd, &h.eft, &h.efdt, &h.xf, &h.xfd, &h.xft,
                                    &h.xfdt};
    for (int k = 0; k < 21; ++k) dst[k]->assign(a[k], a[k] + n);
    h.t.resize(jmax);
    h.d.resize(imax);
    for (int j = 0; j < jmax; ++j) h.t[j] = dexp10(h.logtlo + j * h.logtstp);  // :198-201
    for (int i = 0; i < imax; ++i) h.d[i] = dexp10(h.logdlo + i * h.logdstp);
    h.dt_sav.resize(jmax); h.dt2_sav.resize(jmax); h.dti_sav.resize(jmax);
    h.dt2i_sav.resize(jmax); h.dt3i_sav.resize(jmax);
    h.dd_sav.resize(imax); h.dd2_sav.resize(imax); h.ddi_sav.resize(imax);
    h.dd2i_sav.resize(imax); h.dd3i_sav.resize(imax);
    for (int j = 0; j < jmax - 1; ++j) {  // :301-313
        double dth = h.t[j + 1] - h.t[j];
        double dt2 = dth * dth, dti = 1.0 / dth, dt2i = 1.0 / dt2;
        h.dt_sav[j] = dth; h.dt2_sav[j] = dt2; h.dti_sav[j] = dti; h.dt2i_sav[j] = dt2i;
        h.dt3i_sav[j] = dt2i * dti;
    }
    for (int i = 0; i < imax - 1; ++i) {  // :314-326
        double dd = h.d[i + 1] - h.d[i];
        double dd2 = dd * dd, ddi = 1.0 / dd, dd2i = 1.0 / dd2;
        h.dd_sav[i] = dd; h.dd2_sav[i] = dd2; h.ddi_sav[i] = ddi; h.dd2i_sav[i] = dd2i;
        h.dd3i_sav[i] = dd2i * ddi;
    }
}

namespace {
// quintic hermite basis, helm_core.f:222-241
inline double psi0(double z) { return z * z * z * (z * (-6.0 * z + 15.0) - 10.0) + 1.0; }
inline double dpsi0(double z) { return z * z * (z * (-30.0 * z + 60.0) - 30.0); }
inline double ddpsi0(double z) { return z * (z * (-120.0 * z + 180.0) - 60.0); }
inline double dddpsi0(double z) { return z * (-360.0 * z + 360.0) - 60.0; }
inline double psi1(double z) { return z * (z * z * (z * (-3.0 * z + 8.0) - 6.0) + 1.0); }
inline double dpsi1(double z) { return z * z * (z * (-15.0 * z + 32.0) - 18.0) + 1.0; }
inline double ddpsi1(double z) { return z * (z * (-60.0 * z + 96.0) - 36.0); }
inline double dddpsi1(double z) { return z * (-180.0 * z + 192.0) - 36.0; }
inline double psi2(double z) { return 0.5 * z * z * (z * (z * (-z + 3.0) - 3.0) + 1.0); }
inline double dpsi2(double z) { return 0.5 * z * (z * (z * (-5.0 * z + 12.0) - 9.0) + 2.0); }
inline double ddpsi2(double z) { return 0.5 * (z * (z * (-20.0 * z + 36.0) - 18.0) + 2.0); }
inline double dddpsi2(double z) { return 0.5 * (z * (-60.0 * z + 72.0) - 18.0); }
// cubic hermite basis, helm_core.f:268-278
inline double xpsi0(double z) { return z * z * (2.0 * z - 3.0) + 1.0; }
inline double xdpsi0(double z) { return z * (6.0 * z - 6.0); }
inline double xddpsi0(double z) { return 12.0 * z - 6.0; }
inline double xpsi1(double z) { return z * (z * (z - 2.0) + 1.0); }
inline double xdpsi1(double z) { return z * (3.0 * z - 4.0) + 1.0; }
inline double xddpsi1(double z) { return 6.0 * z - 4.0; }

struct W6 { double w0, w1, w2, w0m, w1m, w2m; };
struct W4 { double w0, w1, w0m, w1m; };

// biquintic statement function h5, helm_core.f:244-262 (fi is 1-based in the source)
inline double h5(const double* fi, const W6& t, const W6& d) {
    return fi[0] * d.w0 * t.w0 + fi[1] * d.w0m * t.w0 + fi[2] * d.w0 * t.w0m + fi[3] * d.w0m * t.w0m +
           fi[4] * d.w0 * t.w1 + fi[5] * d.w0m * t.w1 + fi[6] * d.w0 * t.w1m + fi[7] * d.w0m * t.w1m +
           fi[8] * d.w0 * t.w2 + fi[9] * d.w0m * t.w2 + fi[10] * d.w0 * t.w2m + fi[11] * d.w0m * t.w2m +
           fi[12] * d.w1 * t.w0 + fi[13] * d.w1m * t.w0 + fi[14] * d.w1 * t.w0m + fi[15] * d.w1m * t.w0m +
           fi[16] * d.w2 * t.w0 + fi[17] * d.w2m * t.w0 + fi[18] * d.w2 * t.w0m + fi[19] * d.w2m * t.w0m +
           fi[20] * d.w1 * t.w1 + fi[21] * d.w1m * t.w1 + fi[22] * d.w1 * t.w1m + fi[23] * d.w1m * t.w1m +
           fi[24] * d.w2 * t.w1 + fi[25] * d.w2m * t.w1 + fi[26] * d.w2 * t.w1m + fi[27] * d.w2m * t.w1m +
           fi[28] * d.w1 * t.w2 + fi[29] * d.w1m * t.w2 + fi[30] * d.w1 * t.w2m + fi[31] * d.w1m * t.w2m +
           fi[32] * d.w2 * t.w2 + fi[33] * d.w2m * t.w2 + fi[34] * d.w2 * t.w2m + fi[35] * d.w2m * t.w2m;
}
// bicubic statement function h3, helm_core.f:282-290
inline double h3(const double* fi, const W4& t, const W4& d) {
    return fi[0] * d.w0 * t.w0 + fi[1] * d.w0m * t.w0 + fi[2] * d.w0 * t.w0m + fi[3] * d.w0m * t.w0m +
           fi[4] * d.w0 * t.w1 + fi[5] * d.w0m * t.w1 + fi[6] * d.w0 * t.w1m + fi[7] * d.w0m * t.w1m +
           fi[8] * d.w1 * t.w0 + fi[9] * d.w1m * t.w0 + fi[10] * d.w1 * t.w0m + fi[11] * d.w1m * t.w0m +
           fi[12] * d.w1 * t.w1 + fi[13] * d.w1m * t.w1 + fi[14] * d.w1 * t.w1m + fi[15] * d.w1m * t.w1m;
}
}  // namespace

int helmeos2_electron(const HelmTable& h, double T, double logT, double Rho, double logRho,
                      double abar_in, double zbar_in, HelmElectron& r) {
    (void)logRho;
    r = HelmElectron{};
    // helm_core.f:34-79: with logT >= logT1 = 5 the blend factor alfa is 0 (full e+e-).
    if (!(logT >= 5.0)) return -2;  // blended / no-electron regimes are never reached by NScool

    double abar = abar_in, zbar = zbar_in, temp = T, logtemp = logT, den = Rho;
    double ytot1 = 1.0 / abar;
    double ye = ytot1 * zbar;
    double din = ye * den;
    // helm_core.f:345-371 clipping
    if (temp < h.templo) { temp = h.templo; logtemp = dsm::log10(temp); }
    if (din < h.denlo) return -3;  // would switch to skip_elec_pos
    if (temp > h.temphi) { temp = h.temphi; logtemp = h.logthi; }
    if (din > h.denhi) din = h.denhi;

    // helm_electron_positron.dek:5 density derivative
    double dindd = ye;

    // :21-24 hash locate (1-based in source; here 0-based)
    int jat = (int)((logtemp - h.logtlo) * h.logtstpi) + 1;
    jat = std::max(1, std::min(jat, h.jmax - 1));
    int iat = (int)((dsm::log10(din) - h.logdlo) * h.logdstpi) + 1;
    iat = std::max(1, std::min(iat, h.imax - 1));
    int j0 = jat - 1, i0 = iat - 1;
    const int im = h.imax;
    auto at = [&](const std::vector<double>& a, int di, int dj) {
        return a[(size_t)(i0 + di) + (size_t)im * (j0 + dj)];
    };
    auto gather4 = [&](const std::vector<double>& a, double* fi) {
        fi[0] = at(a, 0, 0); fi[1] = at(a, 1, 0); fi[2] = at(a, 0, 1); fi[3] = at(a, 1, 1);
    };
    double fi[36];
    gather4(h.f, fi); gather4(h.ft, fi + 4); gather4(h.ftt, fi + 8); gather4(h.fd, fi + 12);
    gather4(h.fdd, fi + 16); gather4(h.fdt, fi + 20); gather4(h.fddt, fi + 24);
    gather4(h.fdtt, fi + 28); gather4(h.fddtt, fi + 32);

    // :66-69
    double xt = std::max((temp - h.t[j0]) * h.dti_sav[j0], 0.0);
    double xd = std::max((din - h.d[i0]) * h.ddi_sav[i0], 0.0);
    double mxt = 1.0 - xt, mxd = 1.0 - xd;
    const double dt = h.dt_sav[j0], dt2 = h.dt2_sav[j0], dti = h.dti_sav[j0],
                 dt2i = h.dt2i_sav[j0], dt3i = h.dt3i_sav[j0];
    const double dd = h.dd_sav[i0], dd2 = h.dd2_sav[i0], ddi = h.ddi_sav[i0],
                 dd2i = h.dd2i_sav[i0];

    // :72-130 basis functions and derivatives
    W6 st{psi0(xt), psi1(xt) * dt, psi2(xt) * dt2, psi0(mxt), -psi1(mxt) * dt, psi2(mxt) * dt2};
    W6 sd{psi0(xd), psi1(xd) * dd, psi2(xd) * dd2, psi0(mxd), -psi1(mxd) * dd, psi2(mxd) * dd2};
    W6 dst{dpsi0(xt) * dti, dpsi1(xt), dpsi2(xt) * dt, -dpsi0(mxt) * dti, dpsi1(mxt), -dpsi2(mxt) * dt};
    W6 dsd{dpsi0(xd) * ddi, dpsi1(xd), dpsi2(xd) * dd, -dpsi0(mxd) * ddi, dpsi1(mxd), -dpsi2(mxd) * dd};
    W6 ddst{ddpsi0(xt) * dt2i, ddpsi1(xt) * dti, ddpsi2(xt), ddpsi0(mxt) * dt2i, -ddpsi1(mxt) * dti, ddpsi2(mxt)};
    W6 ddsd{ddpsi0(xd) * dd2i, ddpsi1(xd) * ddi, ddpsi2(xd), ddpsi0(mxd) * dd2i, -ddpsi1(mxd) * ddi, ddpsi2(mxd)};
    W6 dddst{dddpsi0(xt) * dt3i, dddpsi1(xt) * dt2i, dddpsi2(xt) * dti,
             -dddpsi0(mxt) * dt3i, dddpsi1(mxt) * dt2i, -dddpsi2(mxt) * dti};

    // :142-190
    double free_ = h5(fi, st, sd);
    double df_d = h5(fi, st, dsd);
    double df_t = h5(fi, dst, sd);
    double df_dd = h5(fi, st, ddsd);
    double df_tt = h5(fi, ddst, sd);
    double df_dt = h5(fi, dst, dsd);
    double df_ttt = h5(fi, dddst, sd);
    double df_dtt = h5(fi, ddst, dsd);
    double df_ddt = h5(fi, dst, ddsd);

    // :199-229 cubic weights
    W4 ct{xpsi0(xt), xpsi1(xt) * dt, xpsi0(mxt), -xpsi1(mxt) * dt};
    W4 cd{xpsi0(xd), xpsi1(xd) * dd, xpsi0(mxd), -xpsi1(mxd) * dd};
    W4 dct{xdpsi0(xt) * dti, xdpsi1(xt), -xdpsi0(mxt) * dti, xdpsi1(mxt)};
    W4 dcd{xdpsi0(xd) * ddi, xdpsi1(xd), -xdpsi0(mxd) * ddi, xdpsi1(mxd)};

    // :232-254 pressure derivative with density
    double gi[16];
    gather4(h.dpdf, gi); gather4(h.dpdft, gi + 4); gather4(h.dpdfd, gi + 8); gather4(h.dpdfdt, gi + 12);
    double dpepdd_in = h3(gi, ct, cd);
    dpepdd_in = std::max(dpepdd_in, 1.0e-30);

    // :271-297 electron chemical potential
    gather4(h.ef, gi); gather4(h.eft, gi + 4); gather4(h.efd, gi + 8); gather4(h.efdt, gi + 12);
    double etaele = h3(gi, ct, cd);
    double x = h3(gi, ct, dcd);
    double detadd = ye * x;
    double detadt = h3(gi, dct, cd);

    // :331-357 number density
    gather4(h.xf, gi); gather4(h.xft, gi + 4); gather4(h.xfd, gi + 8); gather4(h.xfdt, gi + 12);
    double xnefer = h3(gi, ct, cd);
    x = h3(gi, ct, dcd);
    x = std::max(x, 1.0e-30);
    double dxnedd = ye * x;
    double dxnedt = h3(gi, dct, cd);

    // :395-412 pressure
    x = din * din;
    double pele = x * df_d;
    double dpepdt = x * df_dt;
    double dpepdd = dpepdd_in * dindd;
    // :430-436 entropy
    double sele = -df_t * ye;
    double dsepdt = -df_tt * ye;
    // :460-463 energy
    double eele = ye * free_ + temp * sele;
    double deepdt = temp * dsepdt;

    // helm_core.f:83 / helm_store_results.dek:34: total entropy must be positive
    if (!(sele > 0.0)) return -1;

    r.eele = eele; r.pele = pele; r.sele = sele; r.deept = deepdt;
    r.dpepd = dpepdd; r.dpept = dpepdt; r.etaele = etaele;
    r.free_ = free_; r.df_d = df_d; r.df_t = df_t; r.df_dd = df_dd; r.df_tt = df_tt;
    r.df_dt = df_dt; r.df_ttt = df_ttt; r.df_dtt = df_dtt; r.df_ddt = df_ddt;
    r.xnefer = xnefer; r.dxnedd = dxnedd; r.dxnedt = dxnedt; r.detadd = detadd; r.detadt = detadt;
    return 0;
}

// dStar_eos/private/electron.f:10-51
int get_helm_eos_results(const HelmTable& h, double rho, double T, const Composition& c,
                         Thermo& r, double& eta) {
    r = Thermo{0, 0, 0, 0, 0, 0, 0};
    eta = 0.0;
    // pure neutron matter: abs(Yn - 1.0) <= epsilon(1.0) with REAL(4) epsilon
    if (std::fabs(c.Yn - 1.0) <= (double)1.1920929e-07f) return 0;
    double abar = c.A / (1.0 - c.Yn);
    double zbar = c.Z;
    double logT = dsm::log10(T), logRho = dsm::log10(rho);
    HelmElectron he;
    int ierr = helmeos2_electron(h, T, logT, rho, logRho, abar, zbar, he);
    if (ierr != 0) return ierr;
    r.u = he.eele;
    r.p = he.pele;
    r.s = he.sele;
    r.f = r.u - T * r.s;
    r.cv = he.deept;
    r.dpr = rho * he.dpepd;
    r.dpt = T * he.dpept;
    eta = he.etaele;
    return 0;
}

}  // namespace dso
