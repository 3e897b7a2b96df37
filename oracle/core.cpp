// ORACLE (test infrastructure): constants helper, composition moments, MESA-style
// piecewise-monotone cubic interpolation, superfluid Tc lookup, pcy97 atmosphere.
#include <algorithm>
#include <cmath>

#include "dstar_oracle.hpp"

namespace dso {

double electroncharge() { return std::sqrt(cst::electroncharge2); }

// nucchem/public/nucchem_lib.f:47-133
void compute_composition_moments(int n, const double* Zs, const double* As, const double* abunds,
                                 const MomentsOptions& opt, Composition& comp, double& Xsum,
                                 int& ncharged, int* charged_idx, double* Yion) {
    comp = Composition{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // clear_composition :181-188
    std::vector<double> Y(n), Zion(n), Aion(n);
    Xsum = 0.0;
    if (opt.abunds_are_mass_fractions) {
        for (int i = 0; i < n; ++i) Xsum += abunds[i];
        for (int i = 0; i < n; ++i) Y[i] = abunds[i] / As[i];
    } else {
        for (int i = 0; i < n; ++i) Xsum += abunds[i] * As[i];  // dot_product :94
        for (int i = 0; i < n; ++i) Y[i] = abunds[i];
    }
    if (opt.renormalize_mass_fractions)
        for (int i = 0; i < n; ++i) Y[i] = Y[i] / Xsum;

    double Ye = 0.0, Yn = 0.0;
    for (int i = 0; i < n; ++i) Ye += Y[i] * Zs[i];
    for (int i = 0; i < n; ++i)
        if (Zs[i] == 0.0 && As[i] == 1.0) Yn += Y[i];
    comp.Ye = Ye;
    comp.Yn = Yn;

    ncharged = 0;
    for (int i = 0; i < n; ++i) {
        if (Zs[i] > 0.0 && As[i] > 0.0) {
            Zion[ncharged] = Zs[i];
            Aion[ncharged] = As[i];
            Yion[ncharged] = Y[i];
            charged_idx[ncharged] = i;
            ++ncharged;
        }
    }
    const double* Zm;
    const double* Am;
    const double* Ym;
    int nm;
    if (opt.exclude_neutrons) {
        if (comp.Yn > 0.0 && comp.Yn != 1.0)
            for (int i = 0; i < ncharged; ++i) Yion[i] = Yion[i] / (1.0 - comp.Yn);
        Zm = Zion.data(); Am = Aion.data(); Ym = Yion; nm = ncharged;
    } else {
        Zm = Zs; Am = As; Ym = Y.data(); nm = n;
    }
    // do_compute_moments :121-132
    const double fivethird = 5.0 / 3.0, seventhird = 7.0 / 3.0;
    double sA = 0.0;
    for (int i = 0; i < nm; ++i) sA += std::min(1.0, std::max(Ym[i], 1.0e-50));
    comp.A = 1.0 / sA;
    double sZ = 0, sZ53 = 0, sZ2 = 0, sZ73 = 0, sZ52 = 0, sZZ1 = 0, sZ2XoA2 = 0;
    for (int i = 0; i < nm; ++i) sZ += Ym[i] * Zm[i];
    for (int i = 0; i < nm; ++i) sZ53 += Ym[i] * dsm::pow(Zm[i], fivethird);
    for (int i = 0; i < nm; ++i) sZ2 += Ym[i] * (Zm[i] * Zm[i]);
    for (int i = 0; i < nm; ++i) sZ73 += Ym[i] * dsm::pow(Zm[i], seventhird);
    for (int i = 0; i < nm; ++i) sZ52 += Ym[i] * dsm::pow(Zm[i], 2.5);
    for (int i = 0; i < nm; ++i) sZZ1 += Ym[i] * (Zm[i] * dsm::pow(Zm[i] + 1.0, 1.5));
    for (int i = 0; i < nm; ++i) sZ2XoA2 += Zm[i] * Zm[i] * Ym[i] / Am[i];
    comp.Z = sZ * comp.A;
    comp.Z53 = sZ53 * comp.A;
    comp.Z2 = sZ2 * comp.A;
    comp.Z73 = sZ73 * comp.A;
    comp.Z52 = sZ52 * comp.A;
    comp.ZZ1_32 = sZZ1 * comp.A;
    comp.Z2XoA2 = sZ2XoA2;
    comp.Q = comp.Z2 - comp.Z * comp.Z;
}

// MESA interp_1d_lib interp_pm: Steffen (1990, A&A 239, 443) monotone cubic.
// f is (4,n) column-major, f(1,i) holds the values on entry.
void interp_pm(const double* x, int n, double* f) {
    auto F = [&](int k, int i) -> double& { return f[4 * i + k]; };  // 0-based k,i
    if (n < 2) {
        if (n == 1) { F(1, 0) = 0; F(2, 0) = 0; F(3, 0) = 0; }
        return;
    }
    std::vector<double> h(n), smid(n), s(n);
    for (int i = 0; i < n - 1; ++i) h[i] = x[i + 1] - x[i];
    for (int i = 0; i < n - 1; ++i) smid[i] = (F(0, i + 1) - F(0, i)) / h[i];  // eq. (7)
    if (n == 2) {
        s[0] = smid[0];
        s[1] = smid[0];
    } else {
        for (int i = 1; i < n - 1; ++i) {
            double p = (smid[i - 1] * h[i] + smid[i] * h[i - 1]) / (h[i - 1] + h[i]);  // eq. (8)
            double sg = std::copysign(1.0, smid[i - 1]) + std::copysign(1.0, smid[i]);
            s[i] = sg * std::min(std::min(std::fabs(smid[i - 1]), std::fabs(smid[i])),
                                 0.5 * std::fabs(p));  // eq. (11)
        }
        double p1 = smid[0] * (1.0 + h[0] / (h[0] + h[1])) - smid[1] * h[0] / (h[0] + h[1]);  // (25)
        double pn = smid[n - 2] * (1.0 + h[n - 2] / (h[n - 2] + h[n - 3])) -
                    smid[n - 3] * h[n - 2] / (h[n - 2] + h[n - 3]);  // (26)
        if (p1 * smid[0] <= 0.0) s[0] = 0.0;
        else if (std::fabs(p1) > 2.0 * std::fabs(smid[0])) s[0] = 2.0 * smid[0];
        else s[0] = p1;  // (27)
        if (pn * smid[n - 2] <= 0.0) s[n - 1] = 0.0;
        else if (std::fabs(pn) > 2.0 * std::fabs(smid[n - 2])) s[n - 1] = 2.0 * smid[n - 2];
        else s[n - 1] = pn;  // (28)
    }
    for (int i = 0; i < n - 1; ++i) {
        F(1, i) = s[i];
        F(2, i) = (3.0 * smid[i] - 2.0 * s[i] - s[i + 1]) / h[i];      // eq. (3)
        F(3, i) = (s[i] + s[i + 1] - 2.0 * smid[i]) / (h[i] * h[i]);   // eq. (2)
    }
    F(1, n - 1) = s[n - 1];
    F(2, n - 1) = 0.0;
    F(3, n - 1) = 0.0;
}

static int locate(const double* x, int n, double xval) {
    // k (0-based) with x[k] <= xval < x[k+1], clamped to [0, n-2]
    int lo = 0, hi = n - 1;
    if (xval <= x[0]) return 0;
    if (xval >= x[n - 1]) return n - 2;
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (x[mid] <= xval) lo = mid; else hi = mid;
    }
    return lo;
}

void interp_value_and_slope(const double* x, int n, const double* f, double xval, double& val,
                            double& slope) {
    int k = locate(x, n, xval);
    double dx = xval - x[k];
    const double* c = f + 4 * k;
    val = c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
    slope = c[1] + 2.0 * dx * (c[2] + 1.5 * dx * c[3]);
}

double interp_value(const double* x, int n, const double* f, double xval) {
    int k = locate(x, n, xval);
    double dx = xval - x[k];
    const double* c = f + 4 * k;
    return c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
}

// superfluid/private/load_sf.f:5-78 (table -> monotone cubic coefficients)
void sf_load(SfTable& t, int nv, double kFmin, double kFmax, const double* kF, const double* Tc) {
    t.kF.assign(kF, kF + nv);
    t.f.assign(4 * nv, 0.0);
    for (int i = 0; i < nv; ++i) t.f[4 * i] = Tc[i];
    interp_pm(t.kF.data(), nv, t.f.data());
    t.kF_min = kFmin;
    t.kF_max = kFmax;
    t.loaded = true;
}

// superfluid/public/superfluid_lib.f:75-101
static double sf_get_Tc(const SfTables& sf, double kF, int id) {
    const SfTable& t = sf.tab[id];
    if (kF < t.kF_min || kF > t.kF_max) return 0.0;
    double T = interp_value(t.kF.data(), (int)t.kF.size(), t.f.data(), kF);
    return std::max(T, 0.0) * sf.scale[id];
}

// superfluid_lib.f:65-73
void sf_get_results(const SfTables& sf, double kFp, double kFn, double Tc[3]) {
    Tc[0] = Tc[1] = Tc[2] = 0.0;
    if (sf.tab[0].loaded) Tc[0] = sf_get_Tc(sf, kFp, 0);
    if (sf.tab[1].loaded) Tc[1] = sf_get_Tc(sf, kFn, 1);
    if (sf.tab[2].loaded) Tc[2] = sf_get_Tc(sf, kFn, 2);
}

// dStar_atm/private/pcy97.f:8-44
void pcy97_Teff(double grav, double Plight, int n, const double* lgTb, double* lgTeff,
                double* lgflux) {
    double eta = Plight / 1.193e34;
    double g14 = grav * 1.0e-14;
    for (int i = 0; i < n; ++i) {
        double Tb9 = dsm::pow(10.0, lgTb[i]) * 1.0e-9;
        double Tstar = std::sqrt(7.0 * Tb9 * std::sqrt(g14));
        double zeta = Tb9 - 1.0e-3 * Tstar;
        double Te4Fe = g14 * (dsm::pow(R4(7.0) * zeta, 2.25) + dsm::pow(cst::onethird * zeta, 1.25));
        double Teff6_4;
        if (eta > 0) {
            double Te4a = g14 * dsm::pow(18.1 * Tb9, 2.42);
            double a = dsm::pow(Tb9, cst::fivethird) * (1.2 + dsm::pow(5.3e-6 / eta, 0.38));
            Teff6_4 = (a * Te4Fe + Te4a) / (a + 1.0);
        } else {
            Teff6_4 = Te4Fe;
        }
        lgTeff[i] = dsm::log10(1.0e6 * dsm::pow(Teff6_4, 0.25));
        lgflux[i] = dsm::log10(cst::sigma_SB * Teff6_4 * 1.0e24);
    }
}

}  // namespace dso
