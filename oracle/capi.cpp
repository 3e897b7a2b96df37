// ORACLE (test infrastructure): flat C entry points so tests/ and bench.py's cpu_baseline leg
// can drive the restatement through ctypes.  Not part of the product ABI (see include/).
#include <cstdio>
#include <cstring>

#include "dstar_oracle.hpp"
#include "nscool_oracle.hpp"

using namespace dso;

namespace {
HelmTable g_helm;
bool g_helm_loaded = false;
SfTables g_sf;
Composition comp_from(const double* p) {
    return Composition{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10]};
}
}  // namespace

extern "C" {

int dso_constants(double* out, int n) {
    const double v[] = {cst::pi, cst::twopi, cst::threepi, cst::fourpi, cst::onethird, cst::twothird,
                        cst::fivethird, cst::seventhird, cst::threepisquare, cst::boltzmann,
                        cst::avogadro, cst::clight, cst::clight2, cst::amu, cst::hbar,
                        cst::finestructure, electroncharge(), cst::GMsun, cst::Gnewton, cst::Msun,
                        cst::Mneutron, cst::Mproton, cst::Melectron, cst::Mmuon, cst::mev_to_ergs,
                        cst::fm_to_cm, cst::ergs_to_mev, cst::cm_to_fm, cst::a_Bohr, cst::Rydberg,
                        cst::Thomson, cst::hbarc_n, cst::mass_n, cst::length_n, cst::density_n,
                        cst::pressure_n, cst::temperature_n, cst::Mn_n, cst::Mp_n, cst::Me_n,
                        cst::mass_g, cst::potential_g, cst::length_g, cst::density_g,
                        cst::pressure_g, cst::time_g, cst::arad, cst::sigma_SB};
    int m = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n && i < m; ++i) out[i] = v[i];
    return m;
}

int dso_helm_load(const char* path) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return -1;
    char magic[8];
    int32_t dims[2];
    if (std::fread(magic, 1, 8, fp) != 8 || std::memcmp(magic, "HELMTBL1", 8) != 0) { std::fclose(fp); return -2; }
    if (std::fread(dims, sizeof(int32_t), 2, fp) != 2) { std::fclose(fp); return -2; }
    size_t n = (size_t)dims[0] * dims[1];
    std::vector<std::vector<double>> a(21, std::vector<double>(n));
    const double* ptr[21];
    for (int k = 0; k < 21; ++k) {
        if (std::fread(a[k].data(), sizeof(double), n, fp) != n) { std::fclose(fp); return -3; }
        ptr[k] = a[k].data();
    }
    std::fclose(fp);
    helm_setup(g_helm, dims[0], dims[1], ptr);
    g_helm_loaded = true;
    return 0;
}

// raw electron-positron interpolation: out[7] = eele,pele,sele,deept,dpepd,dpept,etaele
int dso_helm_eval(double T, double rho, double abar, double zbar, double* out) {
    HelmElectron r;
    int ierr = helmeos2_electron(g_helm, T, std::log10(T), rho, std::log10(rho), abar, zbar, r);
    out[0] = r.eele; out[1] = r.pele; out[2] = r.sele; out[3] = r.deept; out[4] = r.dpepd;
    out[5] = r.dpept; out[6] = r.etaele;
    return ierr;
}

void dso_sf_load(int type, int nv, double kFmin, double kFmax, const double* kF, const double* Tc) {
    sf_load(g_sf.tab[type], nv, kFmin, kFmax, kF, Tc);
    set_micro_sf_tables(&g_sf);
}
void dso_sf_scale(const double* s) { for (int i = 0; i < 3; ++i) g_sf.scale[i] = s[i]; }
void dso_sf_get_results(double kFp, double kFn, double* Tc) { sf_get_results(g_sf, kFp, kFn, Tc); }

// moments: returns ncharged; comp_out[11]
int dso_moments(int n, const double* Z, const double* A, const double* abunds, int mass_fracs,
                int exclude_neutrons, int renorm, double* comp_out, double* Xsum, int* charged_idx,
                double* Yion) {
    MomentsOptions o;
    o.abunds_are_mass_fractions = mass_fracs;
    o.exclude_neutrons = exclude_neutrons;
    o.renormalize_mass_fractions = renorm;
    Composition c;
    int nch = 0;
    compute_composition_moments(n, Z, A, abunds, o, c, *Xsum, nch, charged_idx, Yion);
    std::memcpy(comp_out, &c, sizeof(c));
    return nch;
}

void dso_interp_pm(const double* x, int n, double* f) { interp_pm(x, n, f); }
void dso_interp_value_and_slope(const double* x, int n, const double* f, int m, const double* xv,
                                double* val, double* slope) {
    for (int i = 0; i < m; ++i) interp_value_and_slope(x, n, f, xv[i], val[i], slope[i]);
}

void dso_pcy97(double grav, double Plight, int n, const double* lgTb, double* lgTeff, double* lgflux) {
    pcy97_Teff(grav, Plight, n, lgTb, lgTeff, lgflux);
}

// bc09 (atm.cpp): tight != 0 selects Bc09Settings::tight(); info (6, n) = rho_ph, P_ph, kappa_ph, nstep, nfcn, n_eos
int dso_bc09_Teff(double grav, double Plight, double Pb, int n, const double* lgTeff, int tight, double* lgTb,
                  double* lgflux, double* info) {
    if (!g_helm_loaded) return -100;
    std::vector<Bc09Knot> kn(n);
    const int ierr = bc09_Teff(g_helm, grav, Plight, Pb, n, lgTeff, lgTb, lgflux,
                               tight ? Bc09Settings::tight() : Bc09Settings::reference(), kn.data());
    if (info)
        for (int i = 0; i < n; ++i) {
            double* o = info + 6 * i;
            o[0] = kn[i].rho_ph; o[1] = kn[i].P_ph; o[2] = kn[i].kappa_ph; o[3] = kn[i].nstep; o[4] = kn[i].nfcn;
            o[5] = kn[i].n_eos;
        }
    return ierr;
}
int dso_bc09_table(double grav, double Plight, double Pb, int nv, int tight, double* lgTb, double* lgTeff4,
                   double* lgflux4) {
    if (!g_helm_loaded) return -100;
    return bc09_table(g_helm, grav, Plight, Pb, nv, lgTb, lgTeff4, lgflux4,
                      tight ? Bc09Settings::tight() : Bc09Settings::reference());
}

// batched EOS: for k in [0,n): rho[k], T[k], ionic[11*k..], Yion[ncharged*k..], Tcs[3*k..];
// chi_in[k] (<0 => default nuclear size); out res[15*k..], phase[k], chi_out[k]
int dso_eval_crust_eos_c(int n, const double* rho, const double* T, const double* ionic, int ncharged,
                       const double* Zion, const double* Aion, const double* Yion, const double* Tcs,
                       const double* chi_in, const double* eos_ctrl, double* res, int* phase,
                       double* chi_out, double* comps);
int dso_eval_crust_eos(int n, const double* rho, const double* T, const double* ionic, int ncharged,
                       const double* Zion, const double* Aion, const double* Yion, const double* Tcs,
                       const double* chi_in, const double* eos_ctrl, double* res, int* phase,
                       double* chi_out) {
    return dso_eval_crust_eos_c(n, rho, T, ionic, ncharged, Zion, Aion, Yion, Tcs, chi_in, eos_ctrl, res, phase, chi_out, nullptr);
}
int dso_eval_crust_eos_c(int n, const double* rho, const double* T, const double* ionic, int ncharged,
                       const double* Zion, const double* Aion, const double* Yion, const double* Tcs,
                       const double* chi_in, const double* eos_ctrl, double* res, int* phase,
                       double* chi_out, double* comps) {
    if (!g_helm_loaded) return -100;
    EosControls rq;
    if (eos_ctrl) { rq.Gamma_melt = eos_ctrl[0]; rq.rsi_melt = eos_ctrl[1]; rq.Ythresh = eos_ctrl[2]; }
    int worst = 0;
    for (int k = 0; k < n; ++k) {
        Composition c = comp_from(ionic + 11 * k);
        double chi = chi_in ? chi_in[k] : use_default_nuclear_size;
        int ph = 0;
        Thermo cm[4];
        int ierr = eval_crust_eos(g_helm, rq, rho[k], T[k], c, ncharged, Zion, Aion,
                                  Yion + (size_t)ncharged * k, Tcs + 3 * k, res + 15 * k, ph, chi, comps ? cm : nullptr);
        if (comps)
            for (int q = 0; q < 4; ++q) {
                double* o = comps + 28 * (size_t)k + 7 * q;  // P,E,S,F,Cv,dPdlnRho,dPdlnT
                o[0] = cm[q].p; o[1] = cm[q].u; o[2] = cm[q].s; o[3] = cm[q].f; o[4] = cm[q].cv; o[5] = cm[q].dpr; o[6] = cm[q].dpt;
            }
        phase[k] = ph;
        chi_out[k] = chi;
        if (ierr != 0) worst = ierr;
    }
    return worst;
}

// out eps[6*k..] = total,pair,photo,plasma,brems,pbf
void dso_crust_neutrino(int n, const double* rho, const double* T, const double* ionic,
                        const double* chi, const double* Tcn, const int* channels, double* eps) {
    bool ch[5];
    for (int i = 0; i < 5; ++i) ch[i] = channels[i] != 0;
    for (int k = 0; k < n; ++k) {
        CrustNeutrino e;
        crust_neutrino_emissivity(rho[k], T[k], comp_from(ionic + 11 * k), chi[k], Tcn[k], ch, e);
        double* o = eps + 6 * k;
        o[0] = e.total; o[1] = e.pair; o[2] = e.photo; o[3] = e.plasma; o[4] = e.brems; o[5] = e.pbf;
    }
}

// cond_ctrl = {include_neutrons, include_sf_phonons, ee_fmla, eQ_fmla, rad_off, rad_on}
// out K[10*k..] = total,ee,ei,eQ,sf,nQ,np,kap,electron_total,neutron_total
void dso_conductivity(int n, const double* rho, const double* T, const double* chi,
                      const double* Gamma, const double* eta, const double* mu_e,
                      const double* ionic, const double* Tns, const double* cond_ctrl, double* K) {
    CondControls rq;
    rq.include_neutrons = cond_ctrl[0] != 0;
    rq.include_superfluid_phonons = cond_ctrl[1] != 0;
    rq.ee_scattering_fmla = (int)cond_ctrl[2];
    rq.eQ_scattering_fmla = (int)cond_ctrl[3];
    rq.rad_full_off_lgrho = cond_ctrl[4];
    rq.rad_full_on_lgrho = cond_ctrl[5];
    for (int k = 0; k < n; ++k) {
        CondComponents c;
        conductivity(rq, rho[k], T[k], chi[k], Gamma[k], eta[k], mu_e[k], comp_from(ionic + 11 * k),
                     Tns[k], c);
        double* o = K + 10 * k;
        o[0] = c.total; o[1] = c.ee; o[2] = c.ei; o[3] = c.eQ; o[4] = c.sf; o[5] = c.nQ; o[6] = c.np;
        o[7] = c.kap; o[8] = c.electron_total; o[9] = c.neutron_total;
    }
}

// ---- NScool hot path (nscool.cpp)
int dso_micro_tables(const MicroTablesIn* in, double* tab_lnT, double* tab_lnCp, double* tab_lnGamma,
                     double* tab_lnEnu, double* tab_lnK, double* rho_bar1) {
    if (!g_helm_loaded) return -100;
    return micro_tables(g_helm, *in, tab_lnT, tab_lnCp, tab_lnGamma, tab_lnEnu, tab_lnK, rho_bar1);
}
int dso_evolve_epoch(const EvolveIn* in, EvolveState* st, EvolveTrace* trace) {
    return evolve_epoch(*in, *st, trace);
}
int dso_evolve_rhs_jac(const EvolveIn* in, const double* lnT, double* f, double* lo, double* di, double* up, double* surf) {
    return evolve_rhs_jac(*in, lnT, f, lo, di, up, surf);
}
int dso_crust_composition(int nv, int nisos, const double* lgP_tab, const double* Ycoef,
                          const double* Ziso, const double* Aiso, int nz, const double* lgP,
                          double* ionic, double* Yion, int* ncharged) {
    return crust_composition_info(nv, nisos, lgP_tab, Ycoef, Ziso, Aiso, nz, lgP, ionic, Yion, ncharged);
}
int dso_rodasp_test(int problem, double tend, double rtol, double atol, double h0, double* y, int* counters) {
    return rodasp_selftest(problem, tend, rtol, atol, h0, y, counters);
}

}  // extern "C"

// dsmath accuracy / bit-parity probes: fn = 0 exp, 1 log, 2 log10, 3 pow, 4 sin, 5 cos, 6 sinh, 7 cosh, 8 atan
extern "C" void dso_dsmath(int fn, int n, const double* x, const double* y, double* out) {
    for (int i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = dsm::exp(x[i]); break;
            case 1: out[i] = dsm::log(x[i]); break;
            case 2: out[i] = dsm::log10(x[i]); break;
            case 3: out[i] = dsm::pow(x[i], y[i]); break;
            case 4: out[i] = dsm::sin(x[i]); break;
            case 5: out[i] = dsm::cos(x[i]); break;
            case 6: out[i] = dsm::sinh(x[i]); break;
            case 7: out[i] = dsm::cosh(x[i]); break;
            default: out[i] = dsm::atan(x[i]); break;
        }
    }
}
