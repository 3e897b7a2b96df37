// ORACLE (test infrastructure): NScool coefficient pass (NScool_create_model.f:289-444),
// crust composition lookup (dStar_crust_lib.f:78-147), RHS / Jacobian / step
// (NScool_evolve.f:11-568) and the stiff integrator.
//
// The integrator is MESA's isolve(rodasp_solver) in the reference; MESA r15140 is not in the
// tree, so this is Hairer & Wanner's RODAS driver (rodas.f: step controller, Gustafsson
// predictive control, defaults FAC1=5, FAC2=1/6, SAFE=0.9, UROUND=1e-16) with G. Steinebach's
// RODASP coefficient set (rodas.f, METH=3) and a banded (ml=mu=1) LU with partial pivoting
// standing in for lapack_decsol.  PARITY UNPINNED: the reference has no golden for the
// integrator or the light curve (NScool/test/chk:13-16); MESA-specific departures (y_min/y_max
// handling) are approximated: a step whose new solution leaves [zmin,zmax] is redone with h/2.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>

#include "nscool_oracle.hpp"

namespace dso {
using namespace cst;

namespace {
const SfTables* g_micro_sf = nullptr;
Composition comp_at(const double* base, int iz) {
    const double* p = base + 11 * iz;
    return Composition{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10]};
}
}  // namespace
void set_micro_sf_tables(const SfTables* sf) { g_micro_sf = sf; }

// ------------------------------------------------------------ coefficient pass
int micro_tables(const HelmTable& h, const MicroTablesIn& in, double* tab_lnT, double* tab_lnCp,
                 double* tab_lnGamma, double* tab_lnEnu, double* tab_lnK, double* rho_bar1) {
    const int nz = in.nz, nt = number_table_pts, nch = in.ncharged;
    EosControls rq;
    rq.Gamma_melt = in.eos_ctrl[0]; rq.rsi_melt = in.eos_ctrl[1]; rq.Ythresh = in.eos_ctrl[2];
    CondControls cq;
    cq.include_neutrons = in.cond_ctrl[0] != 0; cq.include_superfluid_phonons = in.cond_ctrl[1] != 0;
    cq.ee_scattering_fmla = (int)in.cond_ctrl[2]; cq.eQ_scattering_fmla = (int)in.cond_ctrl[3];
    cq.rad_full_off_lgrho = in.cond_ctrl[4]; cq.rad_full_on_lgrho = in.cond_ctrl[5];
    bool ch[5];
    for (int i = 0; i < 5; ++i) ch[i] = in.nu_channels[i] != 0;
    int status = 0;

    // :340-341
    for (int it = 0; it < nt; ++it)
        tab_lnT[it] = lgT_tab_min * ln10 + (lgT_tab_max - lgT_tab_min) * ln10 * (double)it / (double)(nt - 1);

    double rb1 = (in.rho_bar1_override > 0.0) ? in.rho_bar1_override : in.rho_bar[0];
    if (in.do_cells) {
        for (int iz = 0; iz < nz; ++iz) {  // :344-403
            double* lnCp = tab_lnCp + (size_t)4 * nt * iz;
            double* lnGam = tab_lnGamma + (size_t)4 * nt * iz;
            double* lnEnu = tab_lnEnu + (size_t)4 * nt * iz;
            Composition c = comp_at(in.ionic, iz);
            double rho = in.rho[iz];
            double chi = nuclear_volume_fraction(rho, c, default_nuclear_radius);
            double kn = neutron_wavenumber(rho, c, chi);
            double Tc[3];
            if (in.Tc_cell) { Tc[0] = in.Tc_cell[3 * iz]; Tc[1] = in.Tc_cell[3 * iz + 1]; Tc[2] = in.Tc_cell[3 * iz + 2]; }
            else sf_get_results(*g_micro_sf, 0.0, kn, Tc);
            for (int it = 0; it < nt; ++it) {
                double Ttab = dsm::exp(tab_lnT[it]);
                double res[num_eos_results];
                int phase;
                int ierr = eval_crust_eos(h, rq, rho, Ttab, c, nch, in.Zion, in.Aion,
                                          in.Yion + (size_t)nch * iz, Tc, res, phase, chi);
                if (ierr != 0) status = ierr;
                lnCp[4 * it] = dsm::log(res[i_Cp]);
                lnGam[4 * it] = dsm::log(res[i_Gamma]);
                if (iz == 0 && it == 0)  // :367-369
                    rb1 = rho * dsm::pow(in.P_bar[0] / in.P[0], 1.0 / res[i_chiRho]);
                CrustNeutrino eps;
                crust_neutrino_emissivity(rho, Ttab, c, chi, Tc[1], ch, eps);
                lnEnu[4 * it] = dsm::log(eps.total / rho);
            }
            if (in.turn_on_shell_Urca && iz < nz - 1) {  // :379-384
                if (dsm::log10(in.P_bar[iz]) < in.lgP_shell_Urca && dsm::log10(in.P_bar[iz + 1]) > in.lgP_shell_Urca) {
                    for (int it = 0; it < nt; ++it)
                        lnEnu[4 * it] = dsm::log(dsm::exp(lnEnu[4 * it]) +
                                                 1.0e36 * in.shell_Urca_luminosity_coeff *
                                                     ipow(dsm::exp(tab_lnT[it]) / 5.0e8, 5) / in.dm[iz]);
                }
            }
            interp_pm(tab_lnT, nt, lnEnu);
            interp_pm(tab_lnT, nt, lnCp);
            interp_pm(tab_lnT, nt, lnGam);
        }
    }
    *rho_bar1 = rb1;
    if (in.do_faces) {
        for (int iz = 0; iz < nz; ++iz) {  // :406-441
            double* lnK = tab_lnK + (size_t)4 * nt * iz;
            Composition c = comp_at(in.ionic_bar, iz);
            double rho = (iz == 0) ? rb1 : in.rho_bar[iz];
            for (int it = 0; it < nt; ++it) {
                double Ttab = dsm::exp(tab_lnT[it]);
                double chi = nuclear_volume_fraction(rho, c, default_nuclear_radius);
                double kn = neutron_wavenumber(rho, c, chi);
                double Tc[3];
                if (in.Tc_face) { Tc[0] = in.Tc_face[3 * iz]; Tc[1] = in.Tc_face[3 * iz + 1]; Tc[2] = in.Tc_face[3 * iz + 2]; }
                else sf_get_results(*g_micro_sf, 0.0, kn, Tc);
                double res[num_eos_results];
                int phase;
                int ierr = eval_crust_eos(h, rq, rho, Ttab, c, nch, in.Zion, in.Aion,
                                          in.Yion_bar + (size_t)nch * iz, Tc, res, phase, chi);
                if (ierr != 0) status = ierr;
                CondComponents K;
                conductivity(cq, rho, Ttab, chi, res[i_Gamma], res[i_Theta], res[i_mu_e], c, Tc[1], K);
                lnK[4 * it] = dsm::log(K.total);
            }
            interp_pm(tab_lnT, nt, lnK);
        }
    }
    return status;
}

// ------------------------------------------------------ composition lookup
int crust_composition_info(int nv, int nisos, const double* lgP_tab, const double* Ycoef,
                           const double* Ziso, const double* Aiso, int nz, const double* lgP,
                           double* ionic, double* Yion, int* ncharged) {
    std::vector<double> Y(nisos);
    std::vector<int> idx(nisos);
    double lgPmin = lgP_tab[0], lgPmax = lgP_tab[nv - 1];
    int nch = 0;
    for (int i = 0; i < nisos; ++i) if (Ziso[i] > 0.0 && Aiso[i] > 0.0) ++nch;
    for (int iz = 0; iz < nz; ++iz) {
        double lgP_c = std::max(lgP[iz], lgPmin);  // dStar_crust_lib.f:99-100
        lgP_c = std::min(lgP_c, lgPmax);
        for (int i = 0; i < nisos; ++i) {
            Y[i] = interp_value(lgP_tab, nv, Ycoef + (size_t)4 * nv * i, lgP_c);
            if (Y[i] < 0.0) Y[i] = 0.0;  // :137
        }
        MomentsOptions o;
        o.exclude_neutrons = true;
        o.renormalize_mass_fractions = true;
        Composition c;
        double Xsum;
        int n2;
        compute_composition_moments(nisos, Ziso, Aiso, Y.data(), o, c, Xsum, n2, idx.data(),
                                    Yion + (size_t)nch * iz);
        std::memcpy(ionic + 11 * iz, &c, sizeof(c));
    }
    *ncharged = nch;
    return 0;
}

// ------------------------------------------------------------------ RODASP
namespace {
struct RodaspCoef {
    double a21, a31, a32, a41, a42, a43, a51, a52, a53, a54;
    double c21, c31, c32, c41, c42, c43, c51, c52, c53, c54, c61, c62, c63, c64, c65;
    double gamma, c2, c3, c4;
};
// G. Steinebach (1993), "coefficients for RODAS with order 4 for linear parabolic problems"
const RodaspCoef RP = {
    0.3000000000000000e+01, 0.1831036793486759e+01, 0.4955183967433795e+00, 0.2304376582692669e+01,
    -0.5249275245743001e-01, -0.1176798761832782e+01, -0.7170454962423024e+01, -0.4741636671481785e+01,
    -0.1631002631330971e+02, -0.1062004044111401e+01,
    -0.1200000000000000e+02, -0.8791795173947035e+01, -0.2207865586973518e+01, 0.1081793056857153e+02,
    0.6780270611428266e+01, 0.1953485944642410e+02, 0.3419095006749676e+02, 0.1549671153725963e+02,
    0.5474760875964130e+02, 0.1416005392148534e+02, 0.3462605830930532e+02, 0.1530084976114473e+02,
    0.5699955578662667e+02, 0.1840807009793095e+02, -0.5714285714285717e+01,
    0.25, 0.75, 0.21, 0.63};

// tridiagonal LU with partial pivoting (LAPACK dgttrf/dgttrs algorithm == dgbtrf for kl=ku=1)
struct TriLU {
    int n;
    std::vector<double> dl, d, du, du2;
    std::vector<int> ipiv;
    int factor(int n_, const double* sub, const double* diag, const double* sup) {
        n = n_;
        dl.assign(sub, sub + n); d.assign(diag, diag + n); du.assign(sup, sup + n);
        du2.assign(n, 0.0); ipiv.assign(n, 0);
        // convention: dl[i] = A(i+1,i), du[i] = A(i,i+1), i = 0..n-2
        for (int i = 0; i < n - 1; ++i) {
            if (std::fabs(d[i]) >= std::fabs(dl[i])) {
                ipiv[i] = i;
                if (d[i] == 0.0) return i + 1;
                double fact = dl[i] / d[i];
                dl[i] = fact;
                d[i + 1] = d[i + 1] - fact * du[i];
            } else {
                double fact = d[i] / dl[i];
                d[i] = dl[i];
                dl[i] = fact;
                double temp = du[i];
                du[i] = d[i + 1];
                d[i + 1] = temp - fact * d[i + 1];
                if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -fact * du[i + 1]; }
                ipiv[i] = i + 1;
            }
        }
        if (d[n - 1] == 0.0) return n;
        return 0;
    }
    void solve(double* b) const {
        for (int i = 0; i < n - 1; ++i) {
            if (ipiv[i] == i) b[i + 1] = b[i + 1] - dl[i] * b[i];
            else { double temp = b[i]; b[i] = b[i + 1]; b[i + 1] = temp - dl[i] * b[i]; }
        }
        b[n - 1] = b[n - 1] / d[n - 1];
        if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
        for (int i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
    }
};

struct StiffProblem {
    int n;
    // f(x, y) -> dy ; returns ierr
    std::function<int(double, const double*, double*)> fcn;
    // banded jacobian rows: sup[j] = df(j-1)/dy(j) stored at column j (dfdy(1,j)), diag[j] = dfdy(2,j),
    // sub[j] = df(j+1)/dy(j) (dfdy(3,j)); also refreshes f
    std::function<int(double, const double*, double*, double*, double*, double*)> jac;
    // solout(nr, xold, x, y) -> irtrn
    std::function<int(int, double, double, const double*)> solout;
};

// returns idid (1 ok; -1 input, -2 nmax, -3 step too small, -4 singular, -5 solout/fcn abort)
int rodasp(const StiffProblem& pb, double& x, double* y, double xend, double& h, double max_step,
           int nmax, double rtol, double atol, double ymin, double ymax, int counters[7]) {
    const int n = pb.n;
    const double uround = 1.0e-16, fac1 = 5.0, fac2 = 1.0 / 6.0, safe = 0.9;
    int nfcn = 0, njac = 0, nstep = 0, naccpt = 0, nrejct = 0, ndec = 0, nsol = 0;
    std::vector<double> ynew(n), dy1(n), dy(n), ak1(n), ak2(n), ak3(n), ak4(n), ak5(n), ak6(n), tmp(n);
    std::vector<double> jsup(n), jdiag(n), jsub(n), esub(n), ediag(n), esup(n);
    TriLU lu;
    double posneg = std::copysign(1.0, xend - x);
    double hmax = (max_step == 0.0) ? (xend - x) : max_step;
    hmax = std::min(std::fabs(hmax), std::fabs(xend - x));
    if (std::fabs(h) <= 10.0 * uround) h = 1.0e-6;
    h = std::min(std::fabs(h), hmax);
    h = std::copysign(h, posneg);
    bool reject = false, last = false;
    int nsing = 0, idid = 1;
    double hacc = 0.0, erracc = 0.0, hopt = h;
    double xold = x;
    if (pb.solout && pb.solout(naccpt + 1, xold, x, y) < 0) { idid = 2; goto done; }
    for (;;) {
        if (nstep > nmax) { idid = -2; break; }
        if (0.1 * std::fabs(h) <= std::fabs(x) * uround) { idid = -3; break; }
        if (last) { h = hopt; idid = 1; break; }
        hopt = h;
        if ((x + h * 1.0001 - xend) * posneg >= 0.0) { h = xend - x; last = true; }
        // jacobian (the reference's get_jacobian evaluates f itself, NScool_evolve.f:319)
        if (pb.fcn(x, y, dy1.data()) != 0) { idid = -5; break; }
        ++nfcn;
        ++njac;
        if (pb.jac(x, y, dy1.data(), jsup.data(), jdiag.data(), jsub.data()) != 0) { idid = -5; break; }
        bool redo = true;
        while (redo) {
            redo = false;
            double fac = 1.0 / (h * RP.gamma);
            // E = fac*I - J ; tri storage: sub[i] = E(i+1,i) = -jsub[i] (column i), sup[i] = E(i,i+1) = -jsup[i+1]
            for (int i = 0; i < n; ++i) ediag[i] = fac - jdiag[i];
            for (int i = 0; i < n - 1; ++i) { esub[i] = -jsub[i]; esup[i] = -jsup[i + 1]; }
            if (lu.factor(n, esub.data(), ediag.data(), esup.data()) != 0) {
                ++nsing;
                if (nsing >= 5) { idid = -4; goto done; }
                h = h * 0.5; reject = true; last = false; redo = true;
                continue;
            }
            ++ndec;
            const double hc21 = RP.c21 / h, hc31 = RP.c31 / h, hc32 = RP.c32 / h, hc41 = RP.c41 / h,
                         hc42 = RP.c42 / h, hc43 = RP.c43 / h, hc51 = RP.c51 / h, hc52 = RP.c52 / h,
                         hc53 = RP.c53 / h, hc54 = RP.c54 / h, hc61 = RP.c61 / h, hc62 = RP.c62 / h,
                         hc63 = RP.c63 / h, hc64 = RP.c64 / h, hc65 = RP.c65 / h;
            int ferr = 0;
            // stage 1
            for (int i = 0; i < n; ++i) ak1[i] = dy1[i];
            lu.solve(ak1.data());
            for (int i = 0; i < n; ++i) ynew[i] = y[i] + RP.a21 * ak1[i];
            ferr |= pb.fcn(x + RP.c2 * h, ynew.data(), dy.data());
            for (int i = 0; i < n; ++i) { tmp[i] = hc21 * ak1[i]; ak2[i] = dy[i] + tmp[i]; }
            lu.solve(ak2.data());
            for (int i = 0; i < n; ++i) ynew[i] = y[i] + RP.a31 * ak1[i] + RP.a32 * ak2[i];
            ferr |= pb.fcn(x + RP.c3 * h, ynew.data(), dy.data());
            for (int i = 0; i < n; ++i) { tmp[i] = hc31 * ak1[i] + hc32 * ak2[i]; ak3[i] = dy[i] + tmp[i]; }
            lu.solve(ak3.data());
            for (int i = 0; i < n; ++i) ynew[i] = y[i] + RP.a41 * ak1[i] + RP.a42 * ak2[i] + RP.a43 * ak3[i];
            ferr |= pb.fcn(x + RP.c4 * h, ynew.data(), dy.data());
            for (int i = 0; i < n; ++i) { tmp[i] = hc41 * ak1[i] + hc42 * ak2[i] + hc43 * ak3[i]; ak4[i] = dy[i] + tmp[i]; }
            lu.solve(ak4.data());
            for (int i = 0; i < n; ++i)
                ynew[i] = y[i] + RP.a51 * ak1[i] + RP.a52 * ak2[i] + RP.a53 * ak3[i] + RP.a54 * ak4[i];
            ferr |= pb.fcn(x + h, ynew.data(), dy.data());
            for (int i = 0; i < n; ++i) {
                tmp[i] = hc52 * ak2[i] + hc54 * ak4[i] + hc51 * ak1[i] + hc53 * ak3[i];
                ak5[i] = dy[i] + tmp[i];
            }
            lu.solve(ak5.data());
            // embedded solution
            for (int i = 0; i < n; ++i) ynew[i] = ynew[i] + ak5[i];
            ferr |= pb.fcn(x + h, ynew.data(), dy.data());
            for (int i = 0; i < n; ++i) {
                tmp[i] = hc61 * ak1[i] + hc62 * ak2[i] + hc65 * ak5[i] + hc64 * ak4[i] + hc63 * ak3[i];
                ak6[i] = dy[i] + tmp[i];
            }
            lu.solve(ak6.data());
            for (int i = 0; i < n; ++i) ynew[i] = ynew[i] + ak6[i];
            nsol += 6;
            nfcn += 5;
            ++nstep;
            // bounds / fcn failure: redo with half the step (MESA-style retry; see header note)
            bool oob = (ferr != 0);
            if (ymin != ymax)
                for (int i = 0; i < n; ++i)
                    if (!(ynew[i] >= ymin && ynew[i] <= ymax)) { oob = true; break; }
            if (oob) {
                h = h * 0.5; reject = true; last = false;
                if (naccpt >= 1) ++nrejct;
                if (nstep > nmax) { idid = -2; goto done; }
                if (0.1 * std::fabs(h) <= std::fabs(x) * uround) { idid = -3; goto done; }
                redo = true;
                continue;
            }
            // error estimation
            double err = 0.0;
            for (int i = 0; i < n; ++i) {
                double sk = atol + rtol * std::max(std::fabs(y[i]), std::fabs(ynew[i]));
                err += (ak6[i] / sk) * (ak6[i] / sk);
            }
            err = std::sqrt(err / n);
            double facs = std::max(fac2, std::min(fac1, dsm::pow(err, 0.25) / safe));
            double hnew = h / facs;
            if (err <= 1.0) {
                ++naccpt;
                if (naccpt > 1) {  // Gustafsson predictive controller
                    double facgus = (hacc / h) * dsm::pow(err * err / erracc, 0.25) / safe;
                    facgus = std::max(fac2, std::min(fac1, facgus));
                    facs = std::max(facs, facgus);
                    hnew = h / facs;
                }
                hacc = h;
                erracc = std::max(1.0e-2, err);
                for (int i = 0; i < n; ++i) y[i] = ynew[i];
                xold = x;
                x = x + h;
                if (pb.solout && pb.solout(naccpt + 1, xold, x, y) < 0) { idid = 2; goto done; }
                if (std::fabs(hnew) > hmax) hnew = posneg * hmax;
                if (reject) hnew = posneg * std::min(std::fabs(hnew), std::fabs(h));
                reject = false;
                h = hnew;
            } else {
                reject = true;
                last = false;
                h = hnew;
                if (naccpt >= 1) ++nrejct;
                if (nstep > nmax) { idid = -2; goto done; }
                if (0.1 * std::fabs(h) <= std::fabs(x) * uround) { idid = -3; goto done; }
                redo = true;
            }
        }
    }
done:
    counters[0] = nfcn; counters[1] = njac; counters[2] = nstep; counters[3] = naccpt;
    counters[4] = nrejct; counters[5] = ndec; counters[6] = nsol;
    return idid;
}
}  // namespace

// simple problems for checking the RODASP coefficient set (convergence order, stiff accuracy)
int rodasp_selftest(int problem, double tend, double rtol, double atol, double h0, double* y, int* counters) {
    StiffProblem pb;
    if (problem == 0) {  // y' = -y^2 + nonlinear scalar, y(0) = 1 -> y = 1/(1+t)
        pb.n = 1;
        pb.fcn = [](double, const double* yy, double* f) { f[0] = -yy[0] * yy[0]; return 0; };
        pb.jac = [](double, const double* yy, double*, double* sup, double* d, double* sub) {
            sup[0] = 0; sub[0] = 0; d[0] = -2.0 * yy[0]; return 0; };
    } else if (problem == 1) {  // nonlinear heat equation on 64 cells: u_t = (u^2 u_x)_x discretised
        pb.n = 64;
        pb.fcn = [](double, const double* u, double* f) {
            const int n = 64; const double dx = 1.0 / n;
            for (int i = 0; i < n; ++i) {
                double ul = (i > 0) ? u[i - 1] : 1.0, ur = (i < n - 1) ? u[i + 1] : 2.0;
                double kl = 0.5 * (ul * ul + u[i] * u[i]), kr = 0.5 * (ur * ur + u[i] * u[i]);
                f[i] = (kr * (ur - u[i]) - kl * (u[i] - ul)) / (dx * dx);
            }
            return 0; };
        pb.jac = [](double, const double* u, double*, double* sup, double* d, double* sub) {
            const int n = 64; const double dx = 1.0 / n; const double q = 1.0 / (dx * dx);
            for (int i = 0; i < n; ++i) {
                double ul = (i > 0) ? u[i - 1] : 1.0, ur = (i < n - 1) ? u[i + 1] : 2.0;
                double kl = 0.5 * (ul * ul + u[i] * u[i]), kr = 0.5 * (ur * ur + u[i] * u[i]);
                d[i] = q * (u[i] * (ur - u[i]) - kr - u[i] * (u[i] - ul) - kl);
                // df(i)/du(i+1) stored at column i+1 (sup), df(i)/du(i-1) stored at column i-1 (sub)
                if (i < n - 1) sup[i + 1] = q * (ur * (ur - u[i]) + kr);
                if (i > 0) sub[i - 1] = q * (-ul * (u[i] - ul) + kl);
            }
            sup[0] = 0; sub[n - 1] = 0;
            return 0; };
    } else {  // Prothero-Robinson, autonomous form: y1' = -1e6 (y1 - cos y2) - sin y2, y2' = 1 ; y1 = cos t
        pb.n = 2;
        pb.fcn = [](double, const double* yy, double* f) {
            f[0] = -1.0e6 * (yy[0] - dsm::cos(yy[1])) - dsm::sin(yy[1]);
            f[1] = 1.0;
            return 0; };
        pb.jac = [](double, const double* yy, double*, double* sup, double* d, double* sub) {
            d[0] = -1.0e6; d[1] = 0.0;
            sup[0] = 0.0; sup[1] = -1.0e6 * dsm::sin(yy[1]) - dsm::cos(yy[1]);  // df(0)/dy(1), stored at column 1
            sub[0] = 0.0; sub[1] = 0.0;
            return 0; };
    }
    double x = 0.0, h = h0;
    return rodasp(pb, x, y, tend, h, 0.0, 1000000, rtol, atol, 0.0, 0.0, counters);
}

// ------------------------------------------------------------ thermal step
namespace {
struct Star {
    const EvolveIn& in;
    int n;
    std::vector<double> lnT, T, lnT_bar, T_bar, L;
    std::vector<double> lnK, dlnK, Kcond, lnCp, dlnCp, Cp, lnGam, dlnGam, Gam, lnenu, dlnenu, enu, enuc;
    double Lsurf = 0, dlnLsdlnT = 0, Teff = 0;
    explicit Star(const EvolveIn& i) : in(i), n(i.nz) {
        for (auto* v : {&lnT, &T, &lnT_bar, &T_bar, &L, &lnK, &dlnK, &Kcond, &lnCp, &dlnCp, &Cp, &lnGam,
                        &dlnGam, &Gam, &lnenu, &dlnenu, &enu, &enuc})
            v->assign(n, 0.0);
    }
    // NScool_evolve.f:526-568
    void nuclear_heating() {
        std::fill(enuc.begin(), enuc.end(), 0.0);
        auto do_one = [&](double lgPtop, double lgPbot, double Q) {
            double Ptop = dexp10(lgPtop), Pbot = dexp10(lgPbot);
            double M_norm = 0.0;
            for (int i = 0; i < n; ++i) if (in.P[i] >= Ptop && in.P[i] <= Pbot) M_norm += in.dm[i];
            for (int i = 0; i < n; ++i)
                if (in.P[i] >= Ptop && in.P[i] <= Pbot) enuc[i] = in.Mdot * Q * mev_to_ergs * avogadro / M_norm;
        };
        do_one(in.heat[0], in.heat[1], in.heat[2]);
        do_one(in.heat[3], in.heat[4], in.heat[5]);
        if (in.turn_on_extra_heating) do_one(in.heat[6], in.heat[7], in.heat[8]);
        if (in.enuc_override) for (int i = 0; i < n; ++i) enuc[i] = in.enuc_override[i];
    }
    // :452-459
    void interpolate_temps() {
        T_bar[0] = T[0];
        for (int k = 1; k < n; ++k)
            T_bar[k] = 0.5 * (T[k - 1] * in.dm[k] + T[k] * in.dm[k - 1]) / in.dm_bar[k];
        for (int k = 0; k < n; ++k) lnT_bar[k] = dsm::log(T_bar[k]);
    }
    // :461-524
    void get_coefficients() {
        const int nt = in.n_tab;
        for (int iz = 0; iz < n; ++iz) {
            size_t off = (size_t)4 * nt * iz;
            interp_value_and_slope(in.tab_lnT, nt, in.tab_lnK + off, lnT_bar[iz], lnK[iz], dlnK[iz]);
            interp_value_and_slope(in.tab_lnT, nt, in.tab_lnCp + off, lnT[iz], lnCp[iz], dlnCp[iz]);
            interp_value_and_slope(in.tab_lnT, nt, in.tab_lnGamma + off, lnT[iz], lnGam[iz], dlnGam[iz]);
            interp_value_and_slope(in.tab_lnT, nt, in.tab_lnEnu + off, lnT[iz], lnenu[iz], dlnenu[iz]);
            if (in.reference_cost_quirks) {  // :504-507 whole-array exps inside the zone loop
                for (int k = 0; k < n; ++k) {
                    Kcond[k] = dsm::exp(lnK[k]); Cp[k] = dsm::exp(lnCp[k]);
                    Gam[k] = dsm::exp(lnGam[k]); enu[k] = dsm::exp(lnenu[k]);
                }
            }
        }
        for (int k = 0; k < n; ++k) {
            Kcond[k] = dsm::exp(lnK[k]); Cp[k] = dsm::exp(lnCp[k]);
            Gam[k] = dsm::exp(lnGam[k]); enu[k] = dsm::exp(lnenu[k]);
        }
        nuclear_heating();
    }
    // :433-450
    void evaluate_luminosity() {
        double lgTb = lnT_bar[0] / ln10;
        double lo = in.atm_lgTb[0], hi = in.atm_lgTb[in.atm_nv - 1];
        double lgTc = std::min(std::max(lgTb, lo), hi);  // dStar_atm_lib.f:65
        double lgTeff, dlgTeff, lgflux, dlgflux;
        interp_value_and_slope(in.atm_lgTb, in.atm_nv, in.atm_lgTeff, lgTc, lgTeff, dlgTeff);
        interp_value_and_slope(in.atm_lgTb, in.atm_nv, in.atm_lgflux, lgTc, lgflux, dlgflux);
        Lsurf = in.area[0] * dexp10(lgflux);
        dlnLsdlnT = dlgflux;
        Teff = dexp10(lgTeff);
        L[0] = Lsurf;
        for (int k = 1; k < n; ++k)
            L[k] = -(in.area[k] * in.area[k]) * in.rho_bar[k] * Kcond[k] *
                   (in.ePhi[k - 1] * T[k - 1] - in.ePhi[k] * T[k]) / in.ePhi_bar[k] / in.dm_bar[k];
    }
    bool accreting_fixed() const {
        return in.fix_atmosphere_temperature_when_accreting && in.Mdot > 0.0;
    }
    bool insulating() const { return in.make_inner_boundary_insulating && !in.fix_core_temperature; }
    // :237-283
    int derivatives(const double* y, double* f) {
        for (int k = 0; k < n; ++k) { lnT[k] = y[k]; T[k] = dsm::exp(lnT[k]); }
        interpolate_temps();
        get_coefficients();
        evaluate_luminosity();
        for (int k = 0; k < n - 1; ++k)
            f[k] = ((L[k + 1] * in.e2Phi_bar[k + 1] - L[k] * in.e2Phi_bar[k]) / in.dm[k] / in.e2Phi[k] + enuc[k] -
                    enu[k]) * in.ePhi[k] / Cp[k] / T[k];
        int m = n - 1;
        if (insulating())
            f[m] = ((-L[m] * in.e2Phi_bar[m]) / in.dm[m] / in.e2Phi[m] + enuc[m] - enu[m]) * in.ePhi[m] / Cp[m] / T[m];
        else f[m] = 0.0;
        if (accreting_fixed()) f[0] = 0.0;
        return 0;
    }
    // :285-359 ; band rows in column storage: sup[j] = dfdy(1,j), diag[j] = dfdy(2,j), sub[j] = dfdy(3,j)
    int jacobian(const double* y, double* f, double* sup, double* diag, double* sub) {
        derivatives(y, f);
        std::vector<double> CTinv(n), CTdminv(n), ArKdm(n), wplus(n), wminus(n), dLp(n), dLm(n);
        for (int k = 0; k < n; ++k) {
            CTinv[k] = in.ePhi[k] / (Cp[k] * T[k]);
            CTdminv[k] = CTinv[k] / in.e2Phi[k] / in.dm[k];
            ArKdm[k] = (in.area[k] * in.area[k]) * in.rho_bar[k] * Kcond[k] / in.ePhi_bar[k] / in.dm_bar[k];
        }
        for (int k = 0; k < n - 1; ++k) {  // wplus(1:n-1), dLpdlnT(1:n-1)
            wplus[k] = in.dm[k + 1] * T[k] / (in.dm[k + 1] * T[k] + in.dm[k] * T[k + 1]);
            dLp[k] = L[k + 1] * dlnK[k + 1] * wplus[k] - ArKdm[k + 1] * in.ePhi[k] * T[k];
        }
        for (int k = 1; k < n; ++k) {  // wminus(2:n), dLdlnT(2:n)
            wminus[k] = in.dm[k - 1] * T[k] / (in.dm[k] * T[k - 1] + in.dm[k - 1] * T[k]);
            dLm[k] = L[k] * dlnK[k] * wminus[k] + ArKdm[k] * in.ePhi[k] * T[k];
        }
        for (int k = 0; k < n; ++k) { sup[k] = 0.0; diag[k] = 0.0; sub[k] = 0.0; }
        for (int k = 1; k < n; ++k) sup[k] = CTdminv[k - 1] * in.e2Phi_bar[k] * dLm[k];  // dfdy(1,2:n)
        for (int k = 0; k < n; ++k) diag[k] = -f[k] * (1.0 + dlnCp[k]) - CTinv[k] * enu[k] * dlnenu[k];
        for (int k = 1; k < n - 1; ++k)
            diag[k] = diag[k] + CTdminv[k] * (in.e2Phi_bar[k + 1] * dLp[k] - in.e2Phi_bar[k] * dLm[k]);
        diag[0] = diag[0] + CTdminv[0] * (in.e2Phi_bar[1] * dLp[0] - in.e2Phi_bar[0] * L[0] * dlnLsdlnT);
        int m = n - 1;
        if (insulating()) diag[m] = diag[m] + CTdminv[m] * (-in.e2Phi_bar[m] * dLm[m]);
        else diag[m] = 0.0;
        for (int k = 0; k < n - 1; ++k) sub[k] = -CTdminv[k + 1] * in.e2Phi_bar[k + 1] * dLp[k];  // dfdy(3,1:n-1)
        sub[n - 2] = 0.0;
        if (accreting_fixed()) { sup[1] = 0.0; diag[0] = 0.0; sub[0] = 0.0; }
        return 0;
    }
};
}  // namespace

// get_derivatives (:237-283) and get_jacobian (:285-359) at a given ln T, for the function-level tests: f (nz); rows
// lo/di/up (nz) = d f_k / d lnT_{k-1}, d f_k / d lnT_k, d f_k / d lnT_{k+1} (the band storage of :340-356 transposed
// to rows); surf = Lsurf, dlnLs/dlnT, Teff.  in.Mdot, in.heat select the epoch.
int evolve_rhs_jac(const EvolveIn& in, const double* lnT, double* f, double* lo, double* di, double* up, double* surf) {
    const int n = in.nz;
    Star s(in);
    std::vector<double> y(lnT, lnT + n), sup(n), diag(n), sub(n);
    if (s.accreting_fixed()) y[0] = dsm::log(in.atmosphere_temperature_when_accreting);
    s.jacobian(y.data(), f, sup.data(), diag.data(), sub.data());
    for (int k = 0; k < n; ++k) {
        di[k] = diag[k];
        up[k] = (k < n - 1) ? sup[k + 1] : 0.0;   // dfdy(1, k+1) = d f_k / d y_{k+1}
        lo[k] = (k > 0) ? sub[k - 1] : 0.0;       // dfdy(3, k-1) = d f_k / d y_{k-1}
    }
    if (surf) { surf[0] = s.Lsurf; surf[1] = s.dlnLsdlnT; surf[2] = s.Teff; }
    return 0;
}

int evolve_epoch(const EvolveIn& in, EvolveState& st, EvolveTrace* trace) {
    const int n = in.nz;
    Star s(in);
    std::vector<double> z(st.lnT, st.lnT + n);
    if (s.accreting_fixed()) z[0] = dsm::log(in.atmosphere_temperature_when_accreting);  // :69-71
    double rtol = in.integration_tolerance;
    double zminv = z[0];
    for (int k = 1; k < n; ++k) zminv = std::min(zminv, z[k]);
    double atol = in.integration_tolerance * zminv;  // :76
    double zmin = in.min_lg_T * ln10, zmax = in.max_lg_T * ln10;

    StiffProblem pb;
    pb.n = n;
    pb.fcn = [&](double, const double* y, double* f) { return s.derivatives(y, f); };
    pb.jac = [&](double, const double* y, double* f, double* sup, double* d, double* sub) {
        return s.jacobian(y, f, sup, d, sub); };
    if (trace) trace->count = 0;
    pb.solout = [&](int nr, double xold, double x, const double* y) {  // evaluate_timestep :138-235
        if (in.suppress_first_step_output && x == xold) return 0;
        st.model = nr + in.starting_number_for_profile;
        st.tsec = x + in.epoch_start_time;
        st.dt = x - xold;
        for (int k = 0; k < n; ++k) { s.lnT[k] = y[k]; s.T[k] = dsm::exp(s.lnT[k]); }
        s.interpolate_temps();
        s.get_coefficients();
        s.evaluate_luminosity();
        st.Teff = s.Teff; st.Lsurf = s.Lsurf; st.dlnLsdlnT = s.dlnLsdlnT;
        if (trace && trace->count < trace->capacity) {
            int r = trace->count++;
            double Lnu = 0.0, Lnuc = 0.0;
            for (int k = 0; k < n; ++k) Lnu += s.enu[k] * in.dm[k];
            for (int k = 0; k < n; ++k) Lnuc += s.enuc[k] * in.dm[k];
            trace->model[r] = st.model; trace->tsec[r] = st.tsec; trace->dt[r] = st.dt;
            trace->Teff[r] = s.Teff; trace->Lsurf[r] = s.Lsurf; trace->Lnu[r] = Lnu; trace->Lnuc[r] = Lnuc;
            if (trace->lnT_rows) std::memcpy(trace->lnT_rows + (size_t)r * n, y, sizeof(double) * n);
        }
        return 0;
    };
    double t = 0.0, h = st.dt;
    int idid = rodasp(pb, t, z.data(), in.epoch_duration, h, in.maximum_timestep,
                      in.maximum_number_of_models, rtol, atol, zmin, zmax, st.counters);
    for (int k = 0; k < n; ++k) st.lnT[k] = s.lnT[k];  // s% lnT as left by the last solout
    st.idid = idid;
    return idid < 0 ? idid : 0;
}

}  // namespace dso
