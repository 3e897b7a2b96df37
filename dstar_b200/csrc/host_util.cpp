// Host-side numerical utilities of the product; see host_util.hpp.
#include "host_util.hpp"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace dsh_host {

// ---------------------------------------------------------------- monotone cubic
namespace {
inline double sgn1(double v) { return std::signbit(v) ? -1.0 : 1.0; }
// one-sided three-point end slope with Steffen's monotonicity limiter (eqs. 25-28)
inline double end_slope(double s_near, double s_far, double h_near, double h_far) {
    double p = s_near * (1.0 + h_near / (h_near + h_far)) - s_far * h_near / (h_near + h_far);
    if (p * s_near <= 0.0) return 0.0;
    if (std::fabs(p) > 2.0 * std::fabs(s_near)) return 2.0 * s_near;
    return p;
}
inline int bracket(const double* x, int n, double xv) {
    if (xv <= x[0]) return 0;
    if (xv >= x[n - 1]) return n - 2;
    int k = (int)(std::upper_bound(x, x + n, xv) - x) - 1;
    return std::min(std::max(k, 0), n - 2);
}
}  // namespace

void interp_pm(const double* x, int n, double* f) {
    if (n <= 0) return;
    if (n == 1) { f[1] = f[2] = f[3] = 0.0; return; }
    std::vector<double> h(n - 1), sec(n - 1), slope(n);
    for (int i = 0; i + 1 < n; ++i) {
        h[i] = x[i + 1] - x[i];
        sec[i] = (f[4 * (i + 1)] - f[4 * i]) / h[i];
    }
    if (n == 2) {
        slope[0] = slope[1] = sec[0];
    } else {
        for (int i = 1; i + 1 < n; ++i) {
            double p = (sec[i - 1] * h[i] + sec[i] * h[i - 1]) / (h[i - 1] + h[i]);
            double lim = std::min(std::min(std::fabs(sec[i - 1]), std::fabs(sec[i])), 0.5 * std::fabs(p));
            slope[i] = (sgn1(sec[i - 1]) + sgn1(sec[i])) * lim;
        }
        slope[0] = end_slope(sec[0], sec[1], h[0], h[1]);
        slope[n - 1] = end_slope(sec[n - 2], sec[n - 3], h[n - 2], h[n - 3]);
    }
    for (int i = 0; i + 1 < n; ++i) {
        f[4 * i + 1] = slope[i];
        f[4 * i + 2] = (3.0 * sec[i] - 2.0 * slope[i] - slope[i + 1]) / h[i];
        f[4 * i + 3] = (slope[i] + slope[i + 1] - 2.0 * sec[i]) / (h[i] * h[i]);
    }
    f[4 * (n - 1) + 1] = slope[n - 1];
    f[4 * (n - 1) + 2] = 0.0;
    f[4 * (n - 1) + 3] = 0.0;
}

void interp_value_and_slope(const double* x, int n, const double* f, double xv, double& val, double& slope) {
    const int k = bracket(x, n, xv);
    const double dx = xv - x[k];
    const double* c = f + 4 * k;
    val = c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
    slope = c[1] + 2.0 * dx * (c[2] + 1.5 * dx * c[3]);
}

double interp_value(const double* x, int n, const double* f, double xv) {
    const int k = bracket(x, n, xv);
    const double dx = xv - x[k];
    const double* c = f + 4 * k;
    return c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
}

// ---------------------------------------------------------------------- DOP853
#include "dop853_coef_host.inc"

int Dop853::integrate(const Rhs& f, double& x, double* y, double xend, double h, const Solout& solout) {
    const double uround = 2.3e-16, safe = 0.9, facc1 = 1.0 / 0.333, facc2 = 1.0 / 6.0;
    const double posneg = xend >= x ? 1.0 : -1.0;
    const double hmax = std::fabs(max_step == 0.0 ? xend - x : max_step);
    std::vector<std::vector<double>> K(16, std::vector<double>(n));
    std::vector<double> ynew(n), ytmp(n), yold(n), fnew(n);
    std::vector<std::vector<double>> F(7, std::vector<double>(n));
    nfcn = nstep = naccpt = nrejct = 0;
    double facold = 1.0e-4;
    bool last = false, reject = false;
    f(x, y, K[0].data());
    ++nfcn;
    if (h == 0.0) {  // HINIT
        double dnf = 0, dny = 0;
        for (int i = 0; i < n; ++i) {
            double sk = atol + rtol * std::fabs(y[i]);
            dnf += (K[0][i] / sk) * (K[0][i] / sk);
            dny += (y[i] / sk) * (y[i] / sk);
        }
        h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : std::sqrt(dny / dnf) * 0.01;
        h = std::copysign(std::min(h, hmax), posneg);
        for (int i = 0; i < n; ++i) ytmp[i] = y[i] + h * K[0][i];
        f(x + h, ytmp.data(), K[1].data());
        ++nfcn;
        double der2 = 0;
        for (int i = 0; i < n; ++i) {
            double sk = atol + rtol * std::fabs(y[i]);
            der2 += ((K[1][i] - K[0][i]) / sk) * ((K[1][i] - K[0][i]) / sk);
        }
        der2 = std::sqrt(der2) / h;
        double der12 = std::max(std::fabs(der2), std::sqrt(dnf));
        double h1 = der12 <= 1e-15 ? std::max(1e-6, std::fabs(h) * 1e-3) : std::pow(0.01 / der12, 1.0 / 8.0);
        h = std::copysign(std::min(std::min(100.0 * std::fabs(h), h1), hmax), posneg);
    }
    double xold = x;
    Dense none = [](int, double) { return 0.0; };
    if (solout && solout(1, xold, x, y, none) < 0) return 2;
    for (;;) {
        if (nstep > max_steps) return -2;
        if (0.1 * std::fabs(h) <= std::fabs(x) * uround) return -3;
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
        ++nstep;
        for (int s = 1; s < 12; ++s) {
            for (int i = 0; i < n; ++i) {
                double acc = 0;
                for (int j = 0; j < s; ++j) acc += dop_a[s][j] * K[j][i];
                ytmp[i] = y[i] + h * acc;
            }
            f(x + dop_c[s] * h, ytmp.data(), K[s].data());
        }
        nfcn += 11;
        double err = 0, err2 = 0;
        for (int i = 0; i < n; ++i) {
            double acc = 0, e3 = 0, e5 = 0;
            for (int j = 0; j < 12; ++j) { acc += dop_b[j] * K[j][i]; e3 += dop_e3[j] * K[j][i]; e5 += dop_e5[j] * K[j][i]; }
            ynew[i] = y[i] + h * acc;
            double sk = atol + rtol * std::max(std::fabs(y[i]), std::fabs(ynew[i]));
            err2 += (e3 / sk) * (e3 / sk);
            err += (e5 / sk) * (e5 / sk);
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0) deno = 1.0;
        err = std::fabs(h) * err * std::sqrt(1.0 / (n * deno));
        double fac11 = std::pow(err, 0.125);
        double fac = std::max(facc2, std::min(facc1, fac11 / safe));
        double hnew = h / fac;
        if (err <= 1.0) {
            facold = std::max(err, 1e-4);
            (void)facold;
            ++naccpt;
            f(x + h, ynew.data(), fnew.data());
            ++nfcn;
            // dense output (CONTD8): three extra stages
            K[12] = fnew;
            for (int s = 1; s < 4; ++s) {
                for (int i = 0; i < n; ++i) {
                    double acc = 0;
                    for (int j = 0; j < 12 + s; ++j) acc += dop_a_ext[s][j] * K[j][i];
                    ytmp[i] = y[i] + h * acc;
                }
                f(x + dop_c_ext[s] * h, ytmp.data(), K[12 + s].data());
            }
            nfcn += 3;
            for (int i = 0; i < n; ++i) {
                double dy = ynew[i] - y[i];
                F[0][i] = dy;
                F[1][i] = h * K[0][i] - dy;
                F[2][i] = 2.0 * dy - h * (fnew[i] + K[0][i]);
                for (int r = 0; r < 4; ++r) {
                    double acc = 0;
                    for (int j = 0; j < 16; ++j) acc += dop_d[r][j] * K[j][i];
                    F[3 + r][i] = h * acc;
                }
                yold[i] = y[i];
            }
            const double x0 = x, hh = h;
            Dense dense = [&](int i, double s) {
                double th = (s - x0) / hh, th1 = 1.0 - th;
                return yold[i] + th * (F[0][i] + th1 * (F[1][i] + th * (F[2][i] + th1 * (F[3][i] + th * (F[4][i] +
                       th1 * (F[5][i] + th * F[6][i]))))));
            };
            for (int i = 0; i < n; ++i) y[i] = ynew[i];
            K[0] = fnew;
            xold = x;
            x = x + h;
            if (solout && solout(naccpt + 1, xold, x, y, dense) < 0) return 2;
            if (last) return 1;
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
}

// ---------------------------------------------------------------------- root finder
double find_root(const std::function<double(double)>& f, double x0, double dx0, double epsx, double epsy, int imax,
                 int& ierr) {
    ierr = 0;
    double dx = std::fabs(dx0) > 0 ? std::fabs(dx0) : 1e-2;
    double a = x0 - dx, b = x0 + dx;
    double fa = f(a), fb = f(b);
    int it = 0;
    while (fa * fb > 0.0 && it < imax) {  // expand geometrically towards the side with the smaller |f|
        dx *= 1.6;
        if (std::fabs(fa) < std::fabs(fb)) { a -= dx; fa = f(a); } else { b += dx; fb = f(b); }
        ++it;
    }
    if (fa * fb > 0.0) { ierr = -1; return x0; }
    if (fa == 0.0) return a;
    if (fb == 0.0) return b;
    int side = 0;
    double c = a;
    for (it = 0; it < 200; ++it) {
        c = (fa * b - fb * a) / (fa - fb);
        if (!(c > std::min(a, b) && c < std::max(a, b))) c = 0.5 * (a + b);
        double fc = f(c);
        if (std::fabs(fc) < epsy || std::fabs(b - a) < epsx) return c;
        if (fc * fb > 0.0) { b = c; fb = fc; if (side == -1) fa *= 0.5; side = -1; }
        else { a = c; fa = fc; if (side == +1) fb *= 0.5; side = +1; }
    }
    ierr = -2;
    return c;
}

// ---------------------------------------------------------------------- files
bool read_gap_file(const std::string& path, std::vector<double>& kF, std::vector<double>& Tc, double& kFmin,
                   double& kFmax) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    for (int i = 0; i < 3; ++i) if (!std::getline(in, line)) return false;
    int nv = 0;
    if (!(in >> nv >> kFmin >> kFmax) || nv < 2) return false;
    kF.resize(nv); Tc.resize(nv);
    for (int i = 0; i < nv; ++i) if (!(in >> kF[i])) return false;
    for (int i = 0; i < nv; ++i) if (!(in >> Tc[i])) return false;
    return true;
}

bool read_helm_file(const std::string& path, int& imax, int& jmax, std::vector<std::vector<double>>& arrays) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) return false;
    char magic[8];
    int32_t dims[2];
    bool ok = std::fread(magic, 1, 8, fp) == 8 && std::memcmp(magic, "HELMTBL1", 8) == 0 &&
              std::fread(dims, sizeof(int32_t), 2, fp) == 2;
    if (ok) {
        imax = dims[0]; jmax = dims[1];
        size_t n = (size_t)imax * jmax;
        arrays.assign(21, std::vector<double>(n));
        for (int k = 0; k < 21 && ok; ++k) ok = std::fread(arrays[k].data(), sizeof(double), n, fp) == n;
    }
    std::fclose(fp);
    return ok;
}

}  // namespace dsh_host
