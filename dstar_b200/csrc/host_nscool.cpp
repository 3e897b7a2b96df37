// Host-side mirror of NScool_lib (NScool/public/NScool_lib.f:6-91) above the kernel ABI.
//
// What stays on the host (setup, once per model): the &controls namelist (NScool_ctrls_io.f),
// accretion epochs (NScool_epochs.f), the HZ90 composition table (dStar_crust/preprocessor/src/
// HZ90_comp.f), the crust table lgRho(lgP), lgEps(lgP) (dStar_crust_mod.f:74-171; the P->rho inversion
// evaluates the EOS on the DEVICE in batched form), the crust TOV integration
// (NScool_crust_tov.f:110-303) and zone geometry (NScool_create_model.f:138-243), the pcy97
// atmosphere table (dStar_atm/private/pcy97.f), interp_pm of the atmosphere tables (dStar_atm_mod.f:97-99), user
// hooks.  What runs on the device: composition lookups, the bc09 envelope integrations behind the default
// atmosphere table (dStar_atm/private/bc09.f), the coefficient pass, the whole thermal evolution.
#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dstar_nscool.h"
#include "dconst.cuh"
#include "host_util.hpp"

using namespace dsh_host;
namespace k = dsc;

namespace {
// gravitational units (constants_def.f:62-68)
const double mass_g = k::Msun, potential_g = k::clight2;
const double length_g = k::Gnewton * mass_g / potential_g;
const double density_g = mass_g / (length_g * length_g * length_g);
const double pressure_g = density_g * potential_g;

std::string g_data_dir;

// ------------------------------------------------------------------ per-device state
// One slot per GPU this process uses: the kernel-ABI context, the gap tables loaded into it (with their host copies for
// the hook path), the cached batch of the last create_models, the last kernel timings.  dstar_NScool_init opens the
// first slot; dstar_NScool_run_batch opens one per further device and runs one host thread per slot, each thread bound
// to its slot through t_slot.  Everything else (star handles, crust and atmosphere table caches) is shared by the
// threads: handles are disjoint between them, the caches are guarded by g_cache_mu.
struct GapTable { std::vector<double> kF, f; double kmin = 0, kmax = 0; bool loaded = false; };
struct BatchCache {   // device batch shared by create_models -> evolve_models on the same id list
    std::vector<int> ids;
    dstar_batch* b = nullptr;
    void clear() { if (b) dstar_batch_destroy(b); b = nullptr; ids.clear(); }
};
struct DevSlot {
    int device = -1;
    dstar_ctx* ctx = nullptr;
    double ms[3] = {0, 0, 0};
    std::string loaded_gaps[3];
    GapTable gaps[3];
    double sf_scale[3] = {1, 1, 1};
    BatchCache batch;
    int table_sets[2] = {0, 0};
};
std::vector<std::unique_ptr<DevSlot>> g_slots;   // [0] = the device of dstar_NScool_init
thread_local DevSlot* t_slot = nullptr;          // set by the worker threads of dstar_NScool_run_batch
DevSlot g_no_slot;
DevSlot& cur() { return t_slot ? *t_slot : (g_slots.empty() ? g_no_slot : *g_slots[0]); }
std::recursive_mutex g_cache_mu;                 // crust-table and atmosphere-table caches

// ------------------------------------------------------------------ controls
struct Controls {  // NScool/public/NScool_controls.inc, defaults NScool/defaults/controls.defaults
    int write_interval_for_terminal = 1, write_interval_for_terminal_header = 5, write_interval_for_history = 10,
        write_interval_for_profile = 10, starting_number_for_profile = 0;
    bool suppress_first_step_output = false;
    std::string output_directory = "LOGS", which_solver = "rodasp_solver";
    int maximum_number_of_models = 10000;
    double maximum_timestep = 0.0, integration_tolerance = 1.0e-4, min_lg_temperature_integration = 7.0,
           max_lg_temperature_integration = 9.5;
    bool fix_core_temperature = true;
    double core_temperature = 1.0e8;
    bool make_inner_boundary_insulating = false, fix_atmosphere_temperature_when_accreting = true;
    double atmosphere_temperature_when_accreting = 1.0e8;
    bool load_epochs = false;
    std::string epoch_datafile = "accretion_history";
    double epoch_time_scale = 1.0, epoch_Mdot_scale = 1.0;
    int number_epochs = 1;
    std::vector<double> basic_epoch_Mdots = std::vector<double>(64, 0.0);
    std::vector<double> basic_epoch_boundaries = std::vector<double>(65, 0.0);  // (0:64)
    bool use_other_set_Qimp = false, use_other_sf_critical_temperatures = false, use_other_set_heating = false;
    std::vector<double> extra_real_controls = std::vector<double>(32, 0.0);
    std::vector<double> extra_integer_controls = std::vector<double>(32, 0.0);
    std::vector<double> extra_logical_controls = std::vector<double>(32, 0.0);
    double core_mass = 1.6, core_radius = 11.0, lgPcrust_bot = 32.5, lgPcrust_top = 27.0, target_resolution_lnP = 0.1;
    double lgP_min_heating_outer = 26.86, lgP_max_heating_outer = 30.13, Q_heating_outer = 0.3;
    double lgP_min_heating_inner = 30.42, lgP_max_heating_inner = 31.20, Q_heating_inner = 1.5;
    bool turn_on_extra_heating = false;
    double lgP_min_heating_shallow = 27.0, lgP_max_heating_shallow = 28.0, Q_heating_shallow = 3.0;
    std::string which_proton_1S0_gap = "ns", which_neutron_1S0_gap = "gc", which_neutron_3P2_gap = "t72";
    std::vector<double> scale_sf_critical_temperatures = {1.0, 1.0, 1.0};
    bool use_pcy_for_ee_scattering = false, use_page_for_eQ_scattering = false, use_neutron_conductivity = false,
         use_superfluid_phonon_conductivity = false;
    double rad_full_on_lgrho = -1.0, rad_full_off_lgrho = -1.0;
    double eos_gamma_melt_pt = -1.0, eos_rsi_melt_pt = -1.0, eos_nuclide_abundance_threshold = -1.0,
           eos_pasta_transition_in_fm3 = -1.0, eos_cluster_transition_in_fm3 = -1.0;
    bool use_crust_nu_pair = true, use_crust_nu_photo = true, use_crust_nu_plasma = true,
         use_crust_nu_bremsstrahlung = true, use_crust_nu_pbf = true;
    bool turn_on_shell_Urca = false;
    double shell_Urca_luminosity_coeff = 1.0, lgP_shell_Urca = 28.5;
    double lg_atm_light_element_column = 8.0;
    std::string atm_model = "bc09";
    double crust_reference_temperature = 1.0e8;
    std::string crust_composition = "HZ90";
    bool fix_Qimp = false;
    double Qimp = 0.0;
    int trace_capacity = -1;  // dstar_b200 extension: history rows kept on the device (-1: automatic)
    int profile_capacity = -1;  // dstar_b200 extension: profiles kept on the device per star (-1: automatic, 0: none)
    bool share_identical_tables = false;  // dstar_b200 extension: stars of a batch with identical coefficient-pass inputs share one table set
    std::string base_profile_filename = "profile", profile_manifest_filename = "profiles";
    Controls() { basic_epoch_boundaries[1] = 365.0; }
};

std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }
std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
bool parse_real(std::string v, double& out) {
    v = trim(v);
    for (auto& c : v) if (c == 'd' || c == 'D') c = 'e';
    size_t us = v.find("_dp");
    if (us != std::string::npos) v = v.substr(0, us);
    char* end = nullptr;
    out = std::strtod(v.c_str(), &end);
    return end != v.c_str();
}
bool parse_logical(const std::string& v0, bool& out) {
    std::string v = lower(trim(v0));
    if (v == ".true." || v == "t" || v == "true" || v == ".t.") { out = true; return true; }
    if (v == ".false." || v == "f" || v == "false" || v == ".f.") { out = false; return true; }
    return false;
}
// split "a, b, 3*0.0" into values, expanding Fortran repeat counts
std::vector<std::string> split_values(const std::string& s) {
    std::vector<std::string> out;
    std::string cur;
    bool inq = false;
    char q = 0;
    auto flush = [&]() {
        std::string t = trim(cur);
        cur.clear();
        if (t.empty()) return;
        size_t star = t.find('*');
        if (star != std::string::npos && t[0] != '\'' && t[0] != '"') {
            int rep = std::atoi(t.substr(0, star).c_str());
            for (int i = 0; i < rep; ++i) out.push_back(trim(t.substr(star + 1)));
        } else out.push_back(t);
    };
    for (char c : s) {
        if (inq) { cur += c; if (c == q) inq = false; continue; }
        if (c == '\'' || c == '"') { inq = true; q = c; cur += c; continue; }
        if (c == ',') { flush(); continue; }
        cur += c;
    }
    flush();
    return out;
}
std::string unquote(const std::string& s) {
    std::string t = trim(s);
    if (t.size() >= 2 && (t[0] == '\'' || t[0] == '"')) return t.substr(1, t.size() - 2);
    return t;
}

int set_control(Controls& c, const std::string& name_in, const std::string& value) {
    std::string name = lower(trim(name_in));
    int index = -1;  // explicit 1-based (or 0-based for boundaries) subscript
    bool all = false;
    size_t par = name.find('(');
    if (par != std::string::npos) {
        std::string sub = name.substr(par + 1, name.find(')') - par - 1);
        if (trim(sub) == ":") all = true; else index = std::atoi(sub.c_str());
        name = trim(name.substr(0, par));
    }
    std::vector<std::string> vals = split_values(value);
    if (vals.empty()) return -1;
#define REAL(n) if (name == lower(#n)) { return parse_real(vals[0], c.n) ? 0 : -1; }
#define INT(n) if (name == lower(#n)) { double t; if (!parse_real(vals[0], t)) return -1; c.n = (int)t; return 0; }
#define LOG(n) if (name == lower(#n)) { return parse_logical(vals[0], c.n) ? 0 : -1; }
#define STR(n) if (name == lower(#n)) { c.n = trim(unquote(vals[0])); return 0; }
#define ARR(n, base) if (name == lower(#n)) { \
        if (all) { double t; if (!parse_real(vals[0], t)) return -1; for (auto& e : c.n) e = t; return 0; } \
        int start = (index >= 0) ? index - (base) : 0; \
        for (size_t i = 0; i < vals.size() && start + i < c.n.size(); ++i) { \
            double t; bool lg; \
            if (parse_real(vals[i], t)) c.n[start + i] = t; \
            else if (parse_logical(vals[i], lg)) c.n[start + i] = lg ? 1.0 : 0.0; else return -1; } \
        return 0; }
    INT(write_interval_for_terminal) INT(write_interval_for_terminal_header) INT(write_interval_for_history)
    INT(write_interval_for_profile) INT(starting_number_for_profile) LOG(suppress_first_step_output)
    STR(output_directory) STR(base_profile_filename) STR(profile_manifest_filename) STR(which_solver) INT(maximum_number_of_models) REAL(maximum_timestep)
    REAL(integration_tolerance) REAL(min_lg_temperature_integration) REAL(max_lg_temperature_integration)
    LOG(fix_core_temperature) REAL(core_temperature) LOG(make_inner_boundary_insulating)
    LOG(fix_atmosphere_temperature_when_accreting) REAL(atmosphere_temperature_when_accreting)
    LOG(load_epochs) STR(epoch_datafile) REAL(epoch_time_scale) REAL(epoch_Mdot_scale) INT(number_epochs)
    if (name == "basic_epoch_mdots") {
        if (all) { double t; if (!parse_real(vals[0], t)) return -1; for (auto& e : c.basic_epoch_Mdots) e = t; return 0; }
        int start = (index >= 0) ? index - 1 : 0;
        for (size_t i = 0; i < vals.size() && start + i < c.basic_epoch_Mdots.size(); ++i)
            if (!parse_real(vals[i], c.basic_epoch_Mdots[start + i])) return -1;
        return 0;
    }
    if (name == "basic_epoch_boundaries") {
        if (all) { double t; if (!parse_real(vals[0], t)) return -1; for (auto& e : c.basic_epoch_boundaries) e = t; return 0; }
        int start = (index >= 0) ? index : 0;  // declared (0:max)
        for (size_t i = 0; i < vals.size() && start + i < c.basic_epoch_boundaries.size(); ++i)
            if (!parse_real(vals[i], c.basic_epoch_boundaries[start + i])) return -1;
        return 0;
    }
    LOG(use_other_set_Qimp) LOG(use_other_sf_critical_temperatures) LOG(use_other_set_heating)
    ARR(extra_real_controls, 1) ARR(extra_integer_controls, 1) ARR(extra_logical_controls, 1)
    REAL(core_mass) REAL(core_radius) REAL(lgPcrust_bot) REAL(lgPcrust_top) REAL(target_resolution_lnP)
    if (name == "lgp_min_heating_outer") return parse_real(vals[0], c.lgP_min_heating_outer) ? 0 : -1;
    if (name == "lgp_max_heating_outer") return parse_real(vals[0], c.lgP_max_heating_outer) ? 0 : -1;
    if (name == "q_heating_outer") return parse_real(vals[0], c.Q_heating_outer) ? 0 : -1;
    if (name == "lgp_min_heating_inner") return parse_real(vals[0], c.lgP_min_heating_inner) ? 0 : -1;
    if (name == "lgp_max_heating_inner") return parse_real(vals[0], c.lgP_max_heating_inner) ? 0 : -1;
    if (name == "q_heating_inner") return parse_real(vals[0], c.Q_heating_inner) ? 0 : -1;
    LOG(turn_on_extra_heating)
    if (name == "lgp_min_heating_shallow") return parse_real(vals[0], c.lgP_min_heating_shallow) ? 0 : -1;
    if (name == "lgp_max_heating_shallow") return parse_real(vals[0], c.lgP_max_heating_shallow) ? 0 : -1;
    if (name == "q_heating_shallow") return parse_real(vals[0], c.Q_heating_shallow) ? 0 : -1;
    if (name == "which_proton_1s0_gap") { c.which_proton_1S0_gap = trim(unquote(vals[0])); return 0; }
    if (name == "which_neutron_1s0_gap") { c.which_neutron_1S0_gap = trim(unquote(vals[0])); return 0; }
    if (name == "which_neutron_3p2_gap") { c.which_neutron_3P2_gap = trim(unquote(vals[0])); return 0; }
    ARR(scale_sf_critical_temperatures, 1)
    LOG(use_pcy_for_ee_scattering)
    if (name == "use_page_for_eq_scattering") return parse_logical(vals[0], c.use_page_for_eQ_scattering) ? 0 : -1;
    LOG(use_neutron_conductivity) LOG(use_superfluid_phonon_conductivity) REAL(rad_full_on_lgrho)
    REAL(rad_full_off_lgrho) REAL(eos_gamma_melt_pt) REAL(eos_rsi_melt_pt) REAL(eos_nuclide_abundance_threshold)
    REAL(eos_pasta_transition_in_fm3) REAL(eos_cluster_transition_in_fm3)
    LOG(use_crust_nu_pair) LOG(use_crust_nu_photo) LOG(use_crust_nu_plasma) LOG(use_crust_nu_bremsstrahlung)
    LOG(use_crust_nu_pbf)
    if (name == "turn_on_shell_urca") return parse_logical(vals[0], c.turn_on_shell_Urca) ? 0 : -1;
    if (name == "shell_urca_luminosity_coeff") return parse_real(vals[0], c.shell_Urca_luminosity_coeff) ? 0 : -1;
    if (name == "lgp_shell_urca") return parse_real(vals[0], c.lgP_shell_Urca) ? 0 : -1;
    REAL(lg_atm_light_element_column) STR(atm_model) REAL(crust_reference_temperature) STR(crust_composition)
    if (name == "fix_qimp") return parse_logical(vals[0], c.fix_Qimp) ? 0 : -1;
    if (name == "qimp") return parse_real(vals[0], c.Qimp) ? 0 : -1;
    INT(trace_capacity) INT(profile_capacity)
    if (name == "share_identical_tables") return parse_logical(vals[0], c.share_identical_tables) ? 0 : -1;
    // stored-but-unused controls of the reference (load_model_file, model_file, use_core_nu_*)
    if (name == "load_model_file" || name == "model_file" || name.rfind("use_core_nu_", 0) == 0) return 0;
    return -2;
#undef REAL
#undef INT
#undef LOG
#undef STR
#undef ARR
}

// &controls ... / reader (NScool_ctrls_io.f:100-127)
int read_namelist(Controls& c, const std::string& path, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "unable to open inlist " + path; return -1; }
    std::string line, body;
    bool inside = false;
    while (std::getline(in, line)) {
        // strip comments outside quotes
        std::string s;
        bool inq = false; char q = 0;
        for (char ch : line) {
            if (inq) { s += ch; if (ch == q) inq = false; continue; }
            if (ch == '\'' || ch == '"') { inq = true; q = ch; s += ch; continue; }
            if (ch == '!') break;
            s += ch;
        }
        std::string t = trim(s);
        if (!inside) { if (lower(t).rfind("&controls", 0) == 0) inside = true; continue; }
        if (t == "/" || lower(t) == "&end") break;
        body += s + "\n";
    }
    if (!inside) { err = "no &controls namelist in " + path; return -1; }
    // tokenise into name = value chunks: a new assignment starts at an identifier followed by '='
    std::vector<std::pair<std::string, std::string>> items;
    size_t i = 0, n = body.size();
    std::string name, value;
    auto is_ident = [](char ch) { return std::isalnum((unsigned char)ch) || ch == '_' || ch == '(' || ch == ')' || ch == ':' ; };
    while (i < n) {
        // find '=' and walk back over the identifier
        size_t eq = std::string::npos;
        bool inq = false; char q = 0;
        for (size_t j = i; j < n; ++j) {
            char ch = body[j];
            if (inq) { if (ch == q) inq = false; continue; }
            if (ch == '\'' || ch == '"') { inq = true; q = ch; continue; }
            if (ch == '=') { eq = j; break; }
        }
        if (eq == std::string::npos) { value += body.substr(i); break; }
        size_t e = eq;
        while (e > i && std::isspace((unsigned char)body[e - 1])) --e;
        size_t b = e;
        while (b > i && is_ident(body[b - 1])) --b;
        value += body.substr(i, b - i);
        if (!name.empty()) items.push_back({name, value});
        name = body.substr(b, e - b);
        value.clear();
        i = eq + 1;
    }
    if (!name.empty()) items.push_back({name, value});
    for (auto& it : items) {
        std::string v = it.second;
        for (auto& ch : v) if (ch == '\n') ch = ' ';
        v = trim(v);
        while (!v.empty() && v.back() == ',') v.pop_back();
        int rc = set_control(c, it.first, v);
        if (rc != 0) { err = "bad control '" + trim(it.first) + "' = '" + v + "' in " + path; return -1; }
    }
    return 0;
}

// do_load_epochs (NScool_epochs.f:11-171)
int load_epochs(Controls& c, std::vector<double>& boundaries, std::vector<double>& Mdots, std::string& err) {
    std::ifstream in(c.epoch_datafile);
    if (!in) { err = "unable to open " + c.epoch_datafile; return -1; }
    int n_cycles = 1, col_time = 0, col_Mdot = 1;
    std::string line;
    bool table = false;
    std::vector<std::array<double, 2>> rows;
    while (std::getline(in, line)) {
        size_t ex = line.find('!');
        if (ex != std::string::npos) line = line.substr(0, ex);
        std::string t = trim(line);
        if (t.empty()) continue;
        if (!table) {
            std::istringstream ss(t);
            std::string key; ss >> key;
            if (key == "number_cycles") ss >> n_cycles;
            else if (key == "Mdot_scale") { std::string v; ss >> v; parse_real(v, c.epoch_Mdot_scale); }
            else if (key == "time_scale") { std::string v; ss >> v; parse_real(v, c.epoch_time_scale); }
            else if (key == "columns") {
                std::string rest; std::getline(ss, rest); rest = trim(rest);
                if (!rest.empty() && (rest[0] == '\'' || rest[0] == '"')) rest = rest.substr(1);
                if (rest.rfind("Mdot", 0) == 0) { col_Mdot = 0; col_time = 1; } else { col_time = 0; col_Mdot = 1; }
            } else if (key == "epochs") table = true;
        } else {
            std::istringstream ss(t);
            std::string a, b; ss >> a >> b;
            double x, y;
            if (parse_real(a, x) && parse_real(b, y)) rows.push_back({x, y});
        }
    }
    int n = (int)rows.size();
    if (n < 2) { err = "accretion history has fewer than 2 rows"; return -1; }
    for (auto& r : rows) { r[col_time] *= c.epoch_time_scale * k::julian_day; r[col_Mdot] *= c.epoch_Mdot_scale; }
    c.number_epochs = n_cycles * (n - 1);
    boundaries.assign(c.number_epochs + 1, 0.0);
    Mdots.assign(c.number_epochs, 0.0);
    double delta_t = rows[n - 1][col_time] - rows[0][col_time];
    for (int m = 0; m < n_cycles; ++m)
        for (int j = 0; j < n - 1; ++j) {
            boundaries[m * (n - 1) + j] = rows[j][col_time] + m * delta_t;
            Mdots[m * (n - 1) + j] = rows[j][col_Mdot];
        }
    boundaries[c.number_epochs] = rows[n - 1][col_time] + (n_cycles - 1) * delta_t;
    return 0;
}

// ------------------------------------------------------------------ nuclides / compositions
struct Network {
    std::vector<std::string> names;
    std::vector<double> Z, A, mass_excess;  // MeV
    int ncharged() const { int n = 0; for (size_t i = 0; i < Z.size(); ++i) if (Z[i] > 0 && A[i] > 0) ++n; return n; }
};
// The reference's nuclib_db (JINA reaclib winvn) is an LFS stub; mass excesses are synthesised from a
// Bethe-Weizsaecker formula (they only enter lgEps, dStar_crust_mod.f:434-435).
double synthetic_mass_excess(double Z, double A) {
    if (Z == 0 && A == 1) return 8.07131806;
    double N = A - Z;
    double B = 15.75 * A - 17.8 * std::pow(A, 2.0 / 3.0) - 0.711 * Z * Z / std::cbrt(A) - 23.7 * (A - 2 * Z) * (A - 2 * Z) / A;
    int zi = (int)Z, ni = (int)N;
    if (zi % 2 == 0 && ni % 2 == 0) B += 11.18 / std::sqrt(A);
    else if (zi % 2 == 1 && ni % 2 == 1) B -= 11.18 / std::sqrt(A);
    return Z * 7.28897061 + N * 8.07131806 - B;
}
const char* element_names[] = {"n", "h", "he", "li", "be", "b", "c", "n", "o", "f", "ne", "na", "mg", "al", "si", "p", "s",
                               "cl", "ar", "k", "ca", "sc", "ti", "v", "cr", "mn", "fe", "co", "ni", "cu", "zn", "ga", "ge",
                               "as", "se", "br", "kr", "rb", "sr", "y", "zr", "nb", "mo", "tc", "ru", "rh", "pd", "ag", "cd",
                               "in", "sn", "sb", "te", "i", "xe", "cs", "ba"};
bool nuclide_from_name(const std::string& nm, double& Z, double& A) {
    if (nm == "n") { Z = 0; A = 1; return true; }
    if (nm == "p") { Z = 1; A = 1; return true; }
    size_t i = 0;
    while (i < nm.size() && std::isalpha((unsigned char)nm[i])) ++i;
    std::string el = nm.substr(0, i);
    for (int z = 1; z < (int)(sizeof(element_names) / sizeof(char*)); ++z)
        if (el == element_names[z]) { Z = z; A = std::atof(nm.substr(i).c_str()); return A > 0; }
    return false;
}

// composition moments, host copy of nucchem_lib.f:47-133 (exclude_neutrons = .TRUE.)
void moments_excl_neutrons(const Network& net, const double* Yin, bool renorm, double comp[11], double* Yion) {
    const int n = (int)net.Z.size();
    std::vector<double> Y(Yin, Yin + n);
    double Xsum = 0;
    for (int i = 0; i < n; ++i) Xsum += Y[i] * net.A[i];
    if (renorm) for (auto& y : Y) y = y / Xsum;
    double Ye = 0, Yn = 0;
    for (int i = 0; i < n; ++i) Ye += Y[i] * net.Z[i];
    for (int i = 0; i < n; ++i) if (net.Z[i] == 0 && net.A[i] == 1) Yn += Y[i];
    int nc = 0;
    std::vector<double> Zi, Ai;
    for (int i = 0; i < n; ++i) if (net.Z[i] > 0 && net.A[i] > 0) { Yion[nc++] = Y[i]; Zi.push_back(net.Z[i]); Ai.push_back(net.A[i]); }
    if (Yn > 0 && Yn != 1.0) for (int i = 0; i < nc; ++i) Yion[i] = Yion[i] / (1.0 - Yn);
    double sA = 0, sZ = 0, s53 = 0, s2 = 0, s73 = 0, s52 = 0, sz1 = 0, sx = 0;
    for (int i = 0; i < nc; ++i) sA += std::min(1.0, std::max(Yion[i], 1e-50));
    double A = 1.0 / sA;
    for (int i = 0; i < nc; ++i) {
        double y = Yion[i], z = Zi[i];
        sZ += y * z; s53 += y * std::pow(z, 5.0 / 3.0); s2 += y * (z * z); s73 += y * std::pow(z, 7.0 / 3.0);
        s52 += y * std::pow(z, 2.5); sz1 += y * (z * std::pow(z + 1.0, 1.5)); sx += z * z * y / Ai[i];
    }
    comp[0] = A; comp[1] = sZ * A; comp[2] = s53 * A; comp[3] = s2 * A; comp[4] = s73 * A; comp[5] = s52 * A;
    comp[6] = sz1 * A; comp[7] = sx; comp[8] = Ye; comp[9] = Yn; comp[10] = comp[3] - comp[1] * comp[1];
}

// HZ90_comp.f:8-91 : Haensel & Zdunik (1990) composition, 19 isotopes, 17 transitions
Network hz90_network() {
    const char* iso[] = {"n", "mg40", "mg44", "mg48", "si46", "si50", "si54", "s52", "s56", "s60", "ar56", "ar62", "ar66",
                         "ca56", "ca68", "ti56", "ti88", "cr56", "fe56"};
    Network net;
    for (const char* s : iso) {
        double Z, A;
        nuclide_from_name(s, Z, A);
        net.names.push_back(s); net.Z.push_back(Z); net.A.push_back(A); net.mass_excess.push_back(synthetic_mass_excess(Z, A));
    }
    return net;
}
void hz90_table(const Network& net, const std::vector<double>& lgP, std::vector<double>& Y /* (nisos, n) */) {
    const char* layers[] = {"fe56", "cr56", "ti56", "ca56", "ar56", "s52", "si46", "mg40", "ca68", "ar62", "s56", "si50",
                            "mg44", "ar66", "s60", "si54", "mg48", "ti88"};
    const double Pt[] = {7.235e26, 9.569e27, 1.152e29, 4.747e29, 1.361e30, 1.980e30, 2.253e30, 2.637e30, 2.771e30,
                         3.216e30, 3.825e30, 4.699e30, 6.043e30, 7.233e30, 9.238e30, 1.228e31, 1.602e31};
    const double Xn[] = {0.0, 0.0, 0.0, 0.0, 0.0, 4.0 / 56.0, 10.0 / 56.0, 16.0 / 56.0, 22.0 / 56.0, 50.0 / 112.0,
                         56.0 / 112.0, 62.0 / 112.0, 68.0 / 112.0, 158.0 / 224.0, 164.0 / 224.0, 170.0 / 224.0,
                         176.0 / 224.0, 360.0 / 448.0};
    const int nis = (int)net.Z.size(), n = (int)lgP.size();
    Y.assign((size_t)nis * n, 0.0);
    auto idx = [&](const std::string& nm) { for (int i = 0; i < nis; ++i) if (net.names[i] == nm) return i; return -1; };
    const int in = idx("n");
    double lgPt[17];
    for (int i = 0; i < 17; ++i) lgPt[i] = std::log10(Pt[i]);
    for (int j = 0; j < n; ++j) {
        int layer = 17;
        for (int i = 0; i < 17; ++i) if (lgP[j] <= lgPt[i]) { layer = i; break; }
        int ii = idx(layers[layer]);
        Y[(size_t)nis * j + in] = Xn[layer];
        Y[(size_t)nis * j + ii] = (1.0 - Xn[layer]) / net.A[ii];
    }
}

// ------------------------------------------------------------------ abuntime composition files
// dStar_crust/preprocessor: read_abuntime (abuntime.f:18-133), reduce_abuntime (:135-179), extend_abuntime
// (:181-268), driven as in process_abuntime.f:66-76.  Y is (nion, nz) column-major: Y[k + nion * j].
struct AbunTable {
    std::vector<std::string> isos;
    std::vector<double> lgP, Y;
    double T = 0;
    int nz() const { return (int)lgP.size(); }
    int nion() const { return (int)isos.size(); }
};
std::string fixed_field(const std::string& line, size_t pos, size_t len) {
    return pos < line.size() ? line.substr(pos, len) : std::string();
}
int read_abuntime(const std::string& path, double delta_lgP, AbunTable& out, std::string& err) {
    const double gravity = 1.85052e14, mdot = 8.8e4 * 0.3;  // abuntime.f:9-10
    std::ifstream f(path);
    if (!f) { err = "read_abuntime: cannot open " + path; return -1; }
    std::string line;
    if (!std::getline(f, line)) { err = "read_abuntime: empty file"; return -1; }
    const int nion = std::atoi(fixed_field(line, 1, 5).c_str());        // (1x,i5)
    if (nion <= 0) { err = "read_abuntime: bad species count"; return -1; }
    out.isos.assign(nion, "");
    out.lgP.clear(); out.Y.clear();
    std::vector<double> ycur(nion);
    double last_lgP = -100.0, T = 0.0;
    while (std::getline(f, line)) {
        if (trim(line).empty()) continue;
        // (1x,1e18.10,2e11.3,2X,2E18.10): time, T, rho, EFe, EFn
        double time = 0;
        if (!parse_real(trim(fixed_field(line, 1, 18)), time)) { err = "read_abuntime: bad header line: " + line; return -1; }
        parse_real(trim(fixed_field(line, 19, 11)), T);
        const double lgP = std::log10(gravity * mdot * time);
        // (1x,5(a5,1pe10.3)) repeated
        int k = 0;
        while (k < nion) {
            if (!std::getline(f, line)) { err = "read_abuntime: truncated abundance block"; return -1; }
            for (int c = 0; c < 5 && k < nion; ++c, ++k) {
                out.isos[k] = lower(trim(fixed_field(line, 1 + 15 * c, 5)));
                if (!parse_real(trim(fixed_field(line, 6 + 15 * c, 10)), ycur[k])) { err = "read_abuntime: bad abundance: " + line; return -1; }
            }
        }
        std::getline(f, line);  // separator
        if (lgP > last_lgP + delta_lgP) {  // only keep points that advance lgP by more than the increment (:90-95)
            last_lgP = lgP;
            out.lgP.push_back(lgP);
            out.Y.insert(out.Y.end(), ycur.begin(), ycur.end());
        }
    }
    out.T = T * 1.0e9;
    if (out.lgP.empty()) { err = "read_abuntime: no zones accepted"; return -1; }
    return 0;
}
void reduce_abuntime(const AbunTable& in, double min_abundance, AbunTable& out) {
    const int nion = in.nion(), nz = in.nz();
    std::vector<char> keep(nion, 0);
    for (int j = 0; j < nz; ++j) {
        double mx = -1.0e300;
        for (int k = 0; k < nion; ++k) if (in.isos[k] != "n") mx = std::max(mx, in.Y[k + (size_t)nion * j]);
        for (int k = 0; k < nion; ++k) if (in.Y[k + (size_t)nion * j] > min_abundance * mx) keep[k] = 1;
    }
    out = AbunTable();
    out.T = in.T; out.lgP = in.lgP;
    for (int k = 0; k < nion; ++k) if (keep[k]) out.isos.push_back(in.isos[k]);
    const int nnet = out.nion();
    out.Y.resize((size_t)nnet * nz);
    for (int j = 0; j < nz; ++j) {
        int m = 0;
        for (int k = 0; k < nion; ++k) if (keep[k]) out.Y[m++ + (size_t)nnet * j] = in.Y[k + (size_t)nion * j];
    }
}
int extend_abuntime(const AbunTable& in, double lgP_increment, AbunTable& out, std::string& err) {
    const double lgP_blend_width = 0.2, lgPmin = 22.0, lgPmax = 33.5;  // abuntime.f:13-14,185
    const Network hz = hz90_network();
    const int nion = in.nion(), nz = in.nz(), nhz = (int)hz.names.size();
    out = AbunTable();
    out.T = in.T;
    out.isos = in.isos;
    std::vector<int> idx_hz(nhz);
    for (int i = 0; i < nhz; ++i) {
        auto it = std::find(out.isos.begin(), out.isos.end(), hz.names[i]);
        if (it == out.isos.end()) { out.isos.push_back(hz.names[i]); idx_hz[i] = (int)out.isos.size() - 1; }
        else idx_hz[i] = (int)(it - out.isos.begin());
    }
    const int nnet = out.nion();
    const double lo = *std::min_element(in.lgP.begin(), in.lgP.end()), hi = *std::max_element(in.lgP.begin(), in.lgP.end());
    const int nlow = std::max((int)((lo - lgPmin) / lgP_increment), 1), nhigh = std::max((int)((lgPmax - hi) / lgP_increment), 1);
    const int nzout = nlow + nz + nhigh;
    out.lgP.resize(nzout);
    for (int i = 0; i < nlow; ++i) out.lgP[i] = lo - lgP_increment * (double)(nlow - i);
    for (int j = 0; j < nz; ++j) out.lgP[nlow + j] = in.lgP[j];
    for (int i = 1; i <= nhigh; ++i) out.lgP[nlow + nz + i - 1] = hi + lgP_increment * (double)i;
    std::vector<double> Yhz;
    hz90_table(hz, out.lgP, Yhz);
    out.Y.assign((size_t)nnet * nzout, 0.0);
    for (int j = 0; j < nzout; ++j) {
        const int src = std::min(std::max(j - nlow, 0), nz - 1);  // first / last abuntime column replicated outside
        double beta;
        if (out.lgP[j] <= hi) beta = 0.0;
        else if (out.lgP[j] > hi + lgP_blend_width) beta = 1.0;
        else beta = (out.lgP[j] - hi) / lgP_blend_width;
        const double alpha = 1.0 - beta;
        for (int k = 0; k < nion; ++k) out.Y[k + (size_t)nnet * j] = in.Y[k + (size_t)nion * src] * alpha;
        for (int i = 0; i < nhz; ++i) out.Y[idx_hz[i] + (size_t)nnet * j] += Yhz[i + (size_t)nhz * j] * beta;
    }
    (void)err;
    return 0;
}
int process_abuntime(const std::string& path, double lgP_increment, double abundance_threshold, AbunTable& out, std::string& err) {
    AbunTable raw, red;
    int rc = read_abuntime(path, lgP_increment, raw, err);
    if (rc) return rc;
    reduce_abuntime(raw, abundance_threshold, red);
    return extend_abuntime(red, lgP_increment, out, err);
}

// ------------------------------------------------------------------ tables shared between stars

double host_sf_Tc(int id, double kF) {  // superfluid_lib.f:75-101
    const GapTable& t = cur().gaps[id];
    if (!t.loaded || kF < t.kmin || kF > t.kmax) return 0.0;
    return std::max(interp_value(t.kF.data(), (int)t.kF.size(), t.f.data(), kF), 0.0) * cur().sf_scale[id];
}

struct CrustTable {  // dStar_crust_def.f:10-24
    Network net;
    int nv = 0;
    double T = 0;
    std::vector<double> lgP, Ycoef /* (4*nv, nisos) */, lgRho /* (4*nv) */, lgEps;
    std::vector<double> ionic /* (11, nv) */, Yion /* (ncharged, nv) */, mass_excess /* (ncharged) */;   // the table points' composition (tests)
};
std::map<std::string, std::shared_ptr<CrustTable>> g_crust_cache;

struct AtmTable {  // dStar_atm_def.f:11-19
    int nv = 256;
    std::vector<double> lgTb, lgTeff, lgflux;
};
// bc09 tables already generated in this process, keyed by the exact (grav, Plight, Pb) triple.  (The reference
// caches them on disk under a name built from int(100*log10(.)) of the three numbers, dStar_atm_mod.f:171-181, so
// two stars whose gravities differ by < 2 % silently share whichever table was generated first; here every distinct
// triple gets its own table.)
struct AtmKey {
    double g, pl, pb;
    bool operator<(const AtmKey& o) const { return g != o.g ? g < o.g : (pl != o.pl ? pl < o.pl : pb < o.pb); }
};
std::map<AtmKey, std::shared_ptr<AtmTable>> g_atm_cache;

// ------------------------------------------------------------------ one star
struct Star {
    bool in_use = false, setup = false, created = false, evolved = false;
    Controls c;
    dstar_set_Qimp_fn hook_qimp = nullptr;
    dstar_sf_fn hook_sf = nullptr;
    int nz = 0, nisos = 0, ncharged = 0, number_epochs = 0;
    std::vector<double> epoch_Mdots, epoch_boundaries;
    std::shared_ptr<CrustTable> crust;
    AtmTable atm;
    double Lsurf = 0, dlnLsdlnT = 0, Teff = 0, Tcore = 0, Psurf = 0, Pcore = 0, Mcore = 0, Rcore = 0, ePhicore = 0,
           Mtotal = 0, Rtotal = 0, grav = 0, tsec = 0, dt = 0, Plight = 0;
    int model = 0;
    std::map<std::string, std::vector<double>> z;  // zone arrays by name
    std::vector<double> t_monitor, Teff_monitor, Qb_monitor;
    int counters[7] = {0, 0, 0, 0, 0, 0, 0};
    std::vector<int> idid;
    // history rows
    std::vector<int> h_model;
    std::vector<double> h_tsec, h_Mdot, h_Teff, h_Lsurf, h_Lnu, h_Lnuc, h_dt, h_lnTmax, h_PTmax, h_lnTmin, h_PTmin;
    int h_total = 0;  // history rows the run produced (>= kept ones)
    // captured profiles: model number, time, accretion rate and ln T (nz each)
    std::vector<int> p_model;
    std::vector<double> p_tsec, p_Mdot, p_lnT;
    int p_total = 0;  // profiles the run produced (>= kept ones)
    std::string err;
};
constexpr int max_handles = 1 << 16;  // the reference allows 10 (NScool_def.f:159); batches need many more
std::vector<std::unique_ptr<Star>> g_stars;
thread_local std::string g_err;   // per thread: the workers of run_batch report through their own copy

Star* get(int id) {
    if (id < 1 || id > (int)g_stars.size() || !g_stars[id - 1] || !g_stars[id - 1]->in_use) return nullptr;
    return g_stars[id - 1].get();
}

int ensure_gaps(const Controls& c) {
    const std::string want[3] = {c.which_proton_1S0_gap, c.which_neutron_1S0_gap, c.which_neutron_3P2_gap};
    const char* code[3] = {"p1", "n1", "n3"};
    for (int t = 0; t < 3; ++t) {
        if (cur().loaded_gaps[t] != want[t]) {
            std::string path = g_data_dir + "/Tc_data/" + want[t] + "_" + code[t] + ".data";
            GapTable gt;
            std::vector<double> Tc;
            if (!read_gap_file(path, gt.kF, Tc, gt.kmin, gt.kmax)) { g_err = "unable to load gap table " + path; return -1; }
            gt.f.assign(4 * gt.kF.size(), 0.0);
            for (size_t i = 0; i < gt.kF.size(); ++i) gt.f[4 * i] = Tc[i];
            interp_pm(gt.kF.data(), (int)gt.kF.size(), gt.f.data());
            gt.loaded = true;
            int rc = dstar_ctx_set_gaps(cur().ctx, t, (int)gt.kF.size(), gt.kmin, gt.kmax, gt.kF.data(), gt.f.data());
            if (rc != 0) { g_err = dstar_last_error(); return rc; }
            cur().gaps[t] = gt;
            cur().loaded_gaps[t] = want[t];
        }
        cur().sf_scale[t] = c.scale_sf_critical_temperatures[t];
    }
    return dstar_ctx_set_sf_scale(cur().ctx, cur().sf_scale);
}

// do_generate_crust_table + find_densities (dStar_crust_mod.f:74-171, :282-368); EOS on the device
int build_crust_table(const Controls& c, std::shared_ptr<CrustTable>& out) {
    std::lock_guard<std::recursive_mutex> lock(g_cache_mu);   // one thread builds a table, the others find it cached
    char key[256];
    std::snprintf(key, sizeof key, "%s|%.17g|%s|%s|%.17g|%.17g", c.crust_composition.c_str(),
                  c.crust_reference_temperature, c.which_neutron_1S0_gap.c_str(), c.which_proton_1S0_gap.c_str(),
                  c.scale_sf_critical_temperatures[1], c.eos_gamma_melt_pt);
    auto it = g_crust_cache.find(key);
    if (it != g_crust_cache.end()) { out = it->second; return 0; }
    auto tab = std::make_shared<CrustTable>();
    std::vector<double> Y;
    tab->T = c.crust_reference_temperature;
    if (c.crust_composition == "HZ90") {
        tab->net = hz90_network();
        const int n_hz = 2048;  // generate_HZ90.f:12-14
        tab->lgP.resize(n_hz);
        for (int i = 0; i < n_hz; ++i) tab->lgP[i] = 22.0 + (double)i * (33.5 - 22.0) / (double)(n_hz - 1);
        hz90_table(tab->net, tab->lgP, Y);
    } else {
        // any other name is an abuntime-format composition file <data_dir>/crust_data/<name>[.data] (the
        // reference caches the preprocessed table as <name>.bin; here the preprocessing runs at load time)
        std::string path = g_data_dir + "/crust_data/" + c.crust_composition;
        if (!std::ifstream(path)) path += ".data";
        AbunTable ab;
        std::string err;
        if (process_abuntime(path, 0.005, 0.01, ab, err) != 0) {  // defaults of abuntime.f:11-12
            g_err = "crust_composition '" + c.crust_composition + "': " + err;
            return -1;
        }
        for (const auto& nm : ab.isos) {
            double Z, A;
            if (!nuclide_from_name(nm, Z, A)) { g_err = "crust_composition: unknown nuclide '" + nm + "'"; return -1; }
            tab->net.names.push_back(nm); tab->net.Z.push_back(Z); tab->net.A.push_back(A);
            tab->net.mass_excess.push_back(synthetic_mass_excess(Z, A));
        }
        tab->lgP = ab.lgP;
        Y = ab.Y;
    }
    const int N = (int)tab->lgP.size(), nis = (int)tab->net.Z.size();
    tab->nv = N;
    const int nch = tab->net.ncharged();
    std::vector<double> ionic((size_t)11 * N), Yion((size_t)nch * N), Zion, Aion, ME;
    for (int i = 0; i < nis; ++i) if (tab->net.Z[i] > 0) { Zion.push_back(tab->net.Z[i]); Aion.push_back(tab->net.A[i]); ME.push_back(tab->net.mass_excess[i]); }
    for (int i = 0; i < N; ++i) moments_excl_neutrons(tab->net, &Y[(size_t)nis * i], false, &ionic[(size_t)11 * i], &Yion[(size_t)nch * i]);
    // batched Illinois iteration on f(lgRho) = lgP(rho, Tref) - lgP_want, all table points at once
    const double Pfac = 0.25 * std::pow(k::threepisquare, k::onethird) * k::hbar * k::clight * std::pow(k::avogadro, 4.0 * k::onethird);
    std::vector<double> a(N), b(N), fa(N), fb(N), x(N), fx(N), lgEps(N);
    std::vector<int> side(N, 0), done(N, 0);
    double eos_ctrl[3] = {c.eos_gamma_melt_pt > 0 ? c.eos_gamma_melt_pt : 175.0, c.eos_rsi_melt_pt > 0 ? c.eos_rsi_melt_pt : 140.0,
                          c.eos_nuclide_abundance_threshold > 0 ? c.eos_nuclide_abundance_threshold : (double)1.0e-9f};
    auto eval = [&](const std::vector<double>& lgRho, std::vector<double>& f, std::vector<double>& eps) -> int {
        std::vector<double> rho(N), T(N, tab->T), chi(N), Tcs((size_t)3 * N), res((size_t)15 * N), chio(N);
        std::vector<int32_t> phase(N), ie(N);
        for (int i = 0; i < N; ++i) {
            rho[i] = std::pow(10.0, lgRho[i]);
            const double* io = &ionic[(size_t)11 * i];
            double Yn = io[9];
            double rc = std::min(rho[i], 1.0e14);
            chi[i] = k::onethird * k::fourpi * std::pow((double)1.13f * k::fm_to_cm, 3) * (rc * (1.0 - Yn) / k::amu);
            double nn = rho[i] * Yn / k::amu / (1.0 - chi[i]);
            double kn = std::pow(k::threepisquare * nn, k::onethird) / k::cm_to_fm;
            Tcs[3 * i] = host_sf_Tc(0, 0.0); Tcs[3 * i + 1] = host_sf_Tc(1, kn); Tcs[3 * i + 2] = host_sf_Tc(2, kn);
        }
        int rc = dstar_eval_crust_eos(cur().ctx, N, rho.data(), T.data(), ionic.data(), nch, Zion.data(), Aion.data(),
                                      Yion.data(), Tcs.data(), chi.data(), eos_ctrl, res.data(), phase.data(),
                                      chio.data(), ie.data());
        if (rc <= -100) return rc;
        for (int i = 0; i < N; ++i) {
            f[i] = res[(size_t)15 * i] / k::ln10 - tab->lgP[i];
            double Eint = std::exp(res[(size_t)15 * i + 1]);
            double dot = 0;
            for (int j = 0; j < nch; ++j) dot += Yion[(size_t)nch * i + j] * ME[j];
            eps[i] = std::log10(rho[i] * (1.0 + dot / k::amu_n + Eint / k::clight2));
        }
        return 0;
    };
    std::vector<double> g0(N), tmp(N);
    for (int i = 0; i < N; ++i) {
        g0[i] = std::log10(std::pow(std::pow(10.0, tab->lgP[i]) / Pfac, 0.75) / ionic[(size_t)11 * i + 8]);
        a[i] = g0[i] - 0.05 * g0[i]; b[i] = g0[i] + 0.05 * g0[i];
    }
    int rc = eval(a, fa, tmp); if (rc) { g_err = dstar_last_error(); return rc; }
    rc = eval(b, fb, tmp); if (rc) { g_err = dstar_last_error(); return rc; }
    for (int it = 0; it < 20; ++it) {  // widen brackets that do not straddle the root
        bool any = false;
        for (int i = 0; i < N; ++i) if (fa[i] * fb[i] > 0) { any = true; if (std::fabs(fa[i]) < std::fabs(fb[i])) a[i] -= 0.05 * g0[i]; else b[i] += 0.05 * g0[i]; }
        if (!any) break;
        rc = eval(a, fa, tmp); if (rc) { g_err = dstar_last_error(); return rc; }
        rc = eval(b, fb, tmp); if (rc) { g_err = dstar_last_error(); return rc; }
    }
    for (int i = 0; i < N; ++i) if (!(fa[i] * fb[i] <= 0)) { g_err = "find_densities: unable to bracket root"; return -1; }
    for (int it = 0; it < 100; ++it) {
        for (int i = 0; i < N; ++i) {
            if (done[i]) continue;
            double cpt = (fa[i] * b[i] - fb[i] * a[i]) / (fa[i] - fb[i]);
            if (!(cpt > std::min(a[i], b[i]) && cpt < std::max(a[i], b[i]))) cpt = 0.5 * (a[i] + b[i]);
            x[i] = cpt;
        }
        rc = eval(x, fx, lgEps); if (rc) { g_err = dstar_last_error(); return rc; }
        bool all = true;
        for (int i = 0; i < N; ++i) {
            if (done[i]) continue;
            if (std::fabs(fx[i]) < 1.0e-12 || std::fabs(b[i] - a[i]) < 1.0e-12) { done[i] = 1; continue; }
            all = false;
            if (fx[i] * fb[i] > 0) { b[i] = x[i]; fb[i] = fx[i]; if (side[i] == -1) fa[i] *= 0.5; side[i] = -1; }
            else { a[i] = x[i]; fa[i] = fx[i]; if (side[i] == 1) fb[i] *= 0.5; side[i] = 1; }
        }
        if (all) break;
    }
    rc = eval(x, fx, lgEps); if (rc) { g_err = dstar_last_error(); return rc; }
    // interpolants (dStar_crust_mod.f:141-166)
    tab->Ycoef.assign((size_t)4 * N * nis, 0.0);
    for (int s = 0; s < nis; ++s) {
        double* f = &tab->Ycoef[(size_t)4 * N * s];
        for (int i = 0; i < N; ++i) f[4 * i] = Y[(size_t)nis * i + s];
        interp_pm(tab->lgP.data(), N, f);
    }
    tab->lgRho.assign((size_t)4 * N, 0.0); tab->lgEps.assign((size_t)4 * N, 0.0);
    for (int i = 0; i < N; ++i) { tab->lgRho[4 * i] = x[i]; tab->lgEps[4 * i] = lgEps[i]; }
    interp_pm(tab->lgP.data(), N, tab->lgRho.data());
    interp_pm(tab->lgP.data(), N, tab->lgEps.data());
    tab->ionic = ionic; tab->Yion = Yion; tab->mass_excess = ME;
    g_crust_cache[key] = tab;
    out = tab;
    return 0;
}

// tov_integrate (NScool_crust_tov.f:110-303) + do_setup_crust_zones (NScool_create_model.f:138-243)
int setup_zones(Star& s) {
    Controls& c = s.c;
    s.Mcore = c.core_mass; s.Rcore = c.core_radius; s.Psurf = std::exp(c.lgPcrust_top * k::ln10);
    s.Pcore = std::exp(c.lgPcrust_bot * k::ln10); s.Tcore = c.core_temperature;
    const CrustTable& ct = *s.crust;
    const double lgPmin = ct.lgP.front(), lgPmax = ct.lgP.back();
    const double Mc = s.Mcore, Rc = s.Rcore * 1.0e5 / length_g, rez = c.target_resolution_lnP;
    double y[5] = {0, 0, 0, 0.5 * std::log(1.0 - 2.0 * Mc / Rc), 0};
    double lnP = std::log10(s.Pcore) * k::ln10 - std::log(pressure_g);
    const double lnPend = std::log10(s.Psurf) * k::ln10 - std::log(pressure_g);
    std::vector<double> tp, tr, tb, tm, tphi, tv;
    double last_recorded = lnP + rez;
    Dop853 ode;
    ode.n = 5; ode.rtol = 1.0e-5; ode.atol = 1.0e-6; ode.max_steps = 1000;
    auto rhs = [&](double lnPx, const double* yy, double* dy) {
        double P = std::exp(lnPx);
        double r = yy[0] + Rc, m = yy[2] + Mc;
        double lgP = lnPx / k::ln10 + std::log10(pressure_g);
        double lgPc = std::min(std::max(lgP, lgPmin), lgPmax), lgRho, dlgRho, lgEps, dlgEps;
        interp_value_and_slope(ct.lgP.data(), ct.nv, ct.lgRho.data(), lgPc, lgRho, dlgRho);
        interp_value_and_slope(ct.lgP.data(), ct.nv, ct.lgEps.data(), lgPc, lgEps, dlgEps);
        double eps = std::exp(lgEps * k::ln10), rho = std::exp(lgRho * k::ln10);
        double r2 = r * r, r3 = r2 * r, fourpir2 = k::fourpi * r2;
        double rho_g = rho / density_g, eps_g = eps / density_g;
        double Lambda = 1.0 / std::sqrt(1.0 - 2.0 * m / r);
        double Hfac = 1.0 + P / eps_g, Gfac = 1.0 + k::fourpi * r3 * P / m;
        double g = m / r2 * Gfac * Lambda;
        dy[0] = -P / g / eps_g / Hfac / Lambda;
        dy[1] = -P * fourpir2 / g / Hfac * rho_g / eps_g;
        dy[2] = -P * fourpir2 / g / Hfac / Lambda;
        dy[3] = -P / eps_g / Hfac;
        dy[4] = fourpir2 * Lambda * dy[0];
    };
    auto solout = [&](int, double, double x, const double*, const Dop853::Dense& dense) {
        double xwant = last_recorded - rez;
        while (xwant > x) {
            tp.push_back(std::exp(xwant)); tr.push_back(dense(0, xwant)); tb.push_back(dense(1, xwant));
            tm.push_back(dense(2, xwant)); tphi.push_back(dense(3, xwant)); tv.push_back(dense(4, xwant));
            last_recorded -= rez;
            xwant = last_recorded - rez;
        }
        return 0;
    };
    int idid = ode.integrate(rhs, lnP, y, lnPend, -0.1, solout);
    if (idid < 0) { s.err = "integrating TOV eqns"; return -1; }
    const int nzs = (int)tp.size();
    if (nzs < 4) { s.err = "TOV produced too few zones"; return -1; }
    double phi_corr = 0.5 * std::log(1.0 - 2.0 * (y[2] + Mc) / (y[0] + Rc)) - y[3];
    for (auto& v : tphi) v += phi_corr;
    const int nz = nzs - 1;
    s.nz = nz;
    auto& Z = s.z;
    for (const char* nm : {"dm", "P", "lnT", "T", "ePhi", "e2Phi", "eLambda", "rho", "m", "area", "dm_bar", "P_bar", "lnT_bar",
                           "T_bar", "ePhi_bar", "e2Phi_bar", "eLambda_bar", "rho_bar", "Xneut", "Xneut_bar"})
        Z[nm].assign(nz, 0.0);
    // facial quantities: fortran index i (1..nz) <-> TOV index nzs+1-i (1-based) = nzs-i (0-based ... )
    for (int i = 0; i < nz; ++i) {
        const int t = nzs - 1 - i;  // stov%x(nzs:2:-1)
        Z["P_bar"][i] = tp[t] * pressure_g;
        Z["ePhi_bar"][i] = std::exp(tphi[t]);
        Z["e2Phi_bar"][i] = Z["ePhi_bar"][i] * Z["ePhi_bar"][i];
        Z["m"][i] = tb[t] * mass_g;
        double rr = tr[t] * length_g + s.Rcore * 1.0e5;
        Z["area"][i] = k::fourpi * rr * rr;
        Z["eLambda_bar"][i] = 1.0 / std::sqrt(1.0 - 2.0 * (tm[t] + s.Mcore) / (tr[t] + s.Rcore * 1.0e5 / length_g));
    }
    s.ePhicore = std::exp(tphi[0]);
    for (int i = 0; i < nz - 1; ++i) Z["dm"][i] = Z["m"][i] - Z["m"][i + 1];
    Z["dm"][nz - 1] = Z["m"][nz - 1];
    for (int i = 0; i < nz; ++i) {
        const int t = nzs - 1 - i;
        Z["rho"][i] = Z["dm"][i] / (tv[t] - tv[t - 1]) / (length_g * length_g * length_g);
    }
    for (int i = 0; i < nz - 1; ++i) Z["eLambda"][i] = 0.5 * (Z["eLambda_bar"][i] + Z["eLambda_bar"][i + 1]);
    Z["eLambda"][nz - 1] = 0.5 * (Z["eLambda_bar"][nz - 1] + 1.0 / std::sqrt(1.0 - 2.0 * s.Mcore / (s.Rcore * 1.0e5 / length_g)));
    for (int i = 0; i < nz - 1; ++i) Z["ePhi"][i] = 0.5 * (Z["ePhi_bar"][i] + Z["ePhi_bar"][i + 1]);
    Z["ePhi"][nz - 1] = 0.5 * (Z["ePhi_bar"][nz - 1] + s.ePhicore);
    for (int i = 0; i < nz; ++i) Z["e2Phi"][i] = Z["ePhi"][i] * Z["ePhi"][i];
    for (int i = 0; i < nz - 1; ++i) Z["P"][i] = 0.5 * (Z["P_bar"][i] + Z["P_bar"][i + 1]);
    Z["P"][nz - 1] = 1.5 * Z["P_bar"][nz - 1] - 0.5 * Z["P_bar"][nz - 2];
    for (int i = 0; i < nz; ++i) {
        Z["T"][i] = s.Tcore * s.ePhicore / Z["ePhi"][i];
        Z["T_bar"][i] = s.Tcore * s.ePhicore / Z["ePhi_bar"][i];
        Z["lnT"][i] = std::log(Z["T"][i]);
        Z["lnT_bar"][i] = std::log(Z["T_bar"][i]);
    }
    Z["dm_bar"][0] = 0.5 * Z["dm"][0];
    for (int i = 1; i < nz; ++i) Z["dm_bar"][i] = 0.5 * (Z["dm"][i - 1] + Z["dm"][i]);
    for (int i = 1; i < nz; ++i)
        Z["rho_bar"][i] = 0.5 * (Z["rho"][i - 1] * Z["dm"][i] + Z["rho"][i] * Z["dm"][i - 1]) / Z["dm_bar"][i];
    s.Mtotal = tm[nzs - 1] + s.Mcore;
    s.Rtotal = tr[nzs - 1] * length_g * 1.0e-5 + s.Rcore;
    s.grav = s.Mtotal * mass_g * k::Gnewton / ((s.Rtotal * 1.0e5) * (s.Rtotal * 1.0e5)) * Z["eLambda_bar"][0];
    // atmosphere table (dStar_atm_mod.f:46-106), pcy97.f:8-44
    double Plight = s.grav * std::pow(10.0, c.lg_atm_light_element_column);
    s.Plight = Plight;
    if (c.atm_model == "bc09") { s.atm.nv = 0; return 0; }   // generated for the whole batch at once: build_bc09_tables
    if (c.atm_model != "pcy97") {   // dStar_atm_mod.f:102-104: assertion 'atmosphere prefix is okay'
        s.err = "atm_model '" + c.atm_model + "': expected 'bc09' or 'pcy97'";
        return -1;
    }
    AtmTable& at = s.atm;
    at.nv = 256;
    at.lgTb.resize(at.nv); at.lgTeff.assign(4 * at.nv, 0.0); at.lgflux.assign(4 * at.nv, 0.0);
    const double eta = Plight / 1.193e34, g14 = s.grav * 1.0e-14;
    for (int i = 0; i < at.nv; ++i) {
        at.lgTb[i] = 7.0 + (double)i * (9.5 - 7.0) / (double)(at.nv - 1);
        double Tb9 = std::pow(10.0, at.lgTb[i]) * 1.0e-9;
        double Tstar = std::sqrt(7.0 * Tb9 * std::sqrt(g14));
        double zeta = Tb9 - 1.0e-3 * Tstar;
        double Te4Fe = g14 * (std::pow(7.0 * zeta, 2.25) + std::pow(k::onethird * zeta, 1.25));
        double T4 = Te4Fe;
        if (eta > 0) {
            double Te4a = g14 * std::pow(18.1 * Tb9, 2.42);
            double aa = std::pow(Tb9, k::fivethird) * (1.2 + std::pow(5.3e-6 / eta, 0.38));
            T4 = (aa * Te4Fe + Te4a) / (aa + 1.0);
        }
        at.lgTeff[4 * i] = std::log10(1.0e6 * std::pow(T4, 0.25));
        at.lgflux[4 * i] = std::log10(k::sigma_SB * T4 * 1.0e24);
    }
    interp_pm(at.lgTb.data(), at.nv, at.lgTeff.data());
    interp_pm(at.lgTb.data(), at.nv, at.lgflux.data());
    return 0;
}

// do_generate_atm_table with prefix 'bc09' (dStar_atm_mod.f:46-106) for every star of the batch that asked for it:
// the distinct (grav, Plight, Psurf) triples not yet in the cache go to the device in ONE call (thread = table x
// knot, kernels_atm.cuh), then interp_pm in lgTb on the host.
int build_bc09_tables(const std::vector<Star*>& ss) {
    std::lock_guard<std::recursive_mutex> lock(g_cache_mu);
    std::vector<AtmKey> want;
    for (Star* s : ss) {
        if (s->c.atm_model != "bc09") continue;
        const AtmKey key{s->grav, s->Plight, s->Psurf};
        if (g_atm_cache.count(key) == 0 && std::find_if(want.begin(), want.end(), [&](const AtmKey& q) {
                return !(q < key) && !(key < q); }) == want.end())
            want.push_back(key);
    }
    if (!want.empty()) {
        const int nt = (int)want.size(), nv = 256;   // atm_default_number_table_points, dStar_atm_def.f:4
        std::vector<double> g(nt), pl(nt), pb(nt), lgTeff(nv), lgTb((size_t)nv * nt), lgflux((size_t)nv * nt);
        std::vector<int32_t> ierr(nt);
        for (int t = 0; t < nt; ++t) { g[t] = want[t].g; pl[t] = want[t].pl; pb[t] = want[t].pb; }
        for (int i = 0; i < nv; ++i) lgTeff[i] = 5.5 + (double)i * (7.0 - 5.5) / (double)(nv - 1);   // dStar_atm_def.f:7-8
        const int rc = dstar_atm_bc09_Teff(cur().ctx, nt, g.data(), pl.data(), pb.data(), nv, lgTeff.data(), lgTb.data(),
                                           lgflux.data(), ierr.data(), nullptr);
        if (rc != 0) { g_err = std::string("bc09 atmosphere: ") + dstar_last_error(); return rc; }
        for (int t = 0; t < nt; ++t) {
            auto at = std::make_shared<AtmTable>();
            at->nv = nv;
            at->lgTb.assign(&lgTb[(size_t)nv * t], &lgTb[(size_t)nv * t] + nv);
            at->lgTeff.assign(4 * nv, 0.0); at->lgflux.assign(4 * nv, 0.0);
            for (int i = 0; i < nv; ++i) { at->lgTeff[4 * i] = lgTeff[i]; at->lgflux[4 * i] = lgflux[(size_t)nv * t + i]; }
            // (as in the reference, nothing checks that lgTb(lgTeff) came out increasing: dop853 at rtol 1e-6 takes
            //  steps of several e-folds in P across the opacity ramp and the melting line and an occasional knot is off
            //  by ~1e-2 dex; the lookup's binary search is defined for any array)
            interp_pm(at->lgTb.data(), nv, at->lgTeff.data());
            interp_pm(at->lgTb.data(), nv, at->lgflux.data());
            g_atm_cache[want[t]] = at;
        }
    }
    for (Star* s : ss)
        if (s->c.atm_model == "bc09") s->atm = *g_atm_cache[AtmKey{s->grav, s->Plight, s->Psurf}];
    return 0;
}

dstar_micro_ctrl micro_ctrl_of(const Controls& c) {
    dstar_micro_ctrl m{};
    m.eos_gamma_melt_pt = c.eos_gamma_melt_pt; m.eos_rsi_melt_pt = c.eos_rsi_melt_pt;
    m.eos_nuclide_abundance_threshold = c.eos_nuclide_abundance_threshold;
    m.use_neutron_conductivity = c.use_neutron_conductivity;
    m.use_superfluid_phonon_conductivity = c.use_superfluid_phonon_conductivity;
    m.use_pcy_for_ee_scattering = c.use_pcy_for_ee_scattering; m.use_page_for_eQ_scattering = c.use_page_for_eQ_scattering;
    m.rad_full_off_lgrho = c.rad_full_off_lgrho; m.rad_full_on_lgrho = c.rad_full_on_lgrho;
    m.use_crust_nu_pair = c.use_crust_nu_pair; m.use_crust_nu_photo = c.use_crust_nu_photo;
    m.use_crust_nu_plasma = c.use_crust_nu_plasma; m.use_crust_nu_bremsstrahlung = c.use_crust_nu_bremsstrahlung;
    m.use_crust_nu_pbf = c.use_crust_nu_pbf; m.turn_on_shell_Urca = c.turn_on_shell_Urca;
    m.shell_Urca_luminosity_coeff = c.shell_Urca_luminosity_coeff; m.lgP_shell_Urca = c.lgP_shell_Urca;
    return m;
}
dstar_evolve_ctrl evolve_ctrl_of(const Controls& c) {
    dstar_evolve_ctrl e{};
    e.fix_core_temperature = c.fix_core_temperature; e.make_inner_boundary_insulating = c.make_inner_boundary_insulating;
    e.fix_atmosphere_temperature_when_accreting = c.fix_atmosphere_temperature_when_accreting;
    e.turn_on_extra_heating = c.turn_on_extra_heating; e.suppress_first_step_output = c.suppress_first_step_output;
    e.starting_number_for_profile = c.starting_number_for_profile; e.maximum_number_of_models = c.maximum_number_of_models;
    e.integration_tolerance = c.integration_tolerance; e.min_lg_temperature_integration = c.min_lg_temperature_integration;
    e.max_lg_temperature_integration = c.max_lg_temperature_integration;
    e.maximum_timestep = c.maximum_timestep * k::julian_day;  // NScool_ctrls_io.f:162
    return e;
}

// The coefficient pass and the evolve kernel take these settings once per batch (kernel arguments), and the gap tables
// and their scales live in the context: a batch whose stars disagree on any of them is refused, as it is for the
// nuclide network.  (Per-star quantities -- core temperature, Qimp, heating windows, epochs, atmosphere -- are arrays.)
int check_batch_controls(const std::vector<Star*>& ss, bool micro, bool evolve) {
    const Controls& c0 = ss[0]->c;
    const dstar_micro_ctrl m0 = micro_ctrl_of(c0);
    const dstar_evolve_ctrl e0 = evolve_ctrl_of(c0);
    for (size_t i = 1; i < ss.size(); ++i) {
        const Controls& c = ss[i]->c;
        if (micro) {
            const dstar_micro_ctrl m = micro_ctrl_of(c);
            if (std::memcmp(&m, &m0, sizeof m) != 0) {
                g_err = "stars of one batch must share the EOS, conductivity and neutrino controls (differs in star " + std::to_string(i + 1) + ")";
                return -1;
            }
            if (c.which_proton_1S0_gap != c0.which_proton_1S0_gap || c.which_neutron_1S0_gap != c0.which_neutron_1S0_gap ||
                c.which_neutron_3P2_gap != c0.which_neutron_3P2_gap ||
                c.scale_sf_critical_temperatures != c0.scale_sf_critical_temperatures) {
                g_err = "stars of one batch must share the superfluid gap tables and their scale factors (differs in star " +
                        std::to_string(i + 1) + "); per-star critical temperatures go through other_sf_get_results";
                return -1;
            }
        }
        if (evolve) {
            const dstar_evolve_ctrl e = evolve_ctrl_of(c);
            if (std::memcmp(&e, &e0, sizeof e) != 0) {
                g_err = "stars of one batch must share the integration and boundary controls (differs in star " + std::to_string(i + 1) + ")";
                return -1;
            }
            if (c.trace_capacity != c0.trace_capacity || c.profile_capacity != c0.profile_capacity ||
                c.write_interval_for_profile != c0.write_interval_for_profile) {
                g_err = "stars of one batch must share the history/profile capture settings";
                return -1;
            }
        }
    }
    return 0;
}

int g_evolve_mapping = 0;   // dstar_NScool_set_evolve_mapping


void concat(const std::vector<Star*>& ss, const char* name, std::vector<double>& out) {
    out.clear();
    for (Star* s : ss) { auto& v = s->z[name]; out.insert(out.end(), v.begin(), v.end()); }
}
}  // namespace

extern "C" {

int dstar_host_interp_pm(const double* x, int n, double* f) {
    if (!x || !f || n < 1) return -1;
    interp_pm(x, n, f);
    return 0;
}

int dstar_abuntime_table(const char* path, double lgP_increment, double abundance_threshold, int32_t* nnet, int32_t* nz,
                         char* names, double* Z, double* A, double* lgP, double* Y, int cap_net, int cap_z) {
    if (!path || !nnet || !nz) return -1;
    AbunTable ab;
    std::string err;
    if (process_abuntime(path, lgP_increment, abundance_threshold, ab, err) != 0) { g_err = err; return -1; }
    *nnet = ab.nion(); *nz = ab.nz();
    if (!names && !Z && !A && !lgP && !Y) return 0;  // size query
    if (cap_net < ab.nion() || cap_z < ab.nz()) { g_err = "abuntime_table: buffers too small"; return -2; }
    for (int k = 0; k < ab.nion(); ++k) {
        double z = 0, a = 0;
        if (!nuclide_from_name(ab.isos[k], z, a)) { g_err = "abuntime_table: unknown nuclide '" + ab.isos[k] + "'"; return -3; }
        if (names) { std::memset(names + 6 * k, 0, 6); std::strncpy(names + 6 * k, ab.isos[k].c_str(), 5); }
        if (Z) Z[k] = z;
        if (A) A[k] = a;
    }
    if (lgP) std::copy(ab.lgP.begin(), ab.lgP.end(), lgP);
    if (Y) std::copy(ab.Y.begin(), ab.Y.end(), Y);
    return 0;
}

int dstar_ctx_load_helm_file(dstar_ctx* ctx, const char* path) {
    int imax, jmax;
    std::vector<std::vector<double>> a;
    if (!read_helm_file(path, imax, jmax, a)) return DSTAR_ERR_ARG;
    const double* ptr[21];
    for (int i = 0; i < 21; ++i) ptr[i] = a[i].data();
    return dstar_ctx_set_helm(ctx, imax, jmax, ptr);
}

int dstar_ctx_load_gap_file(dstar_ctx* ctx, int type, const char* path) {
    std::vector<double> kF, Tc;
    double kmin, kmax;
    if (!read_gap_file(path, kF, Tc, kmin, kmax)) return DSTAR_ERR_ARG;
    std::vector<double> f(4 * kF.size(), 0.0);
    for (size_t i = 0; i < kF.size(); ++i) f[4 * i] = Tc[i];
    interp_pm(kF.data(), (int)kF.size(), f.data());
    return dstar_ctx_set_gaps(ctx, type, (int)kF.size(), kmin, kmax, kF.data(), f.data());
}

namespace {
// open the slot of `device` (context + HELM table); slot 0 is the device of dstar_NScool_init
int open_slot(int device, DevSlot** out) {
    for (auto& sl : g_slots) if (sl->device == device) { *out = sl.get(); return 0; }
    std::unique_ptr<DevSlot> sl(new DevSlot);
    sl->device = device;
    int rc = dstar_ctx_create(device, &sl->ctx);
    if (rc != 0) { g_err = dstar_last_error(); return rc; }
    rc = dstar_ctx_load_helm_file(sl->ctx, (g_data_dir + "/helm_table.bin").c_str());
    if (rc != 0) { g_err = "unable to read helm table in " + g_data_dir; dstar_ctx_destroy(sl->ctx); return -1; }
    g_slots.push_back(std::move(sl));
    *out = g_slots.back().get();
    return 0;
}
}  // namespace

int dstar_NScool_init(const char* data_dir, int device) {
    if (!g_slots.empty()) return 0;
    g_data_dir = data_dir ? data_dir : "";
    DevSlot* sl = nullptr;
    return open_slot(device, &sl);
}

void dstar_NScool_shutdown(void) {
    g_stars.clear();
    g_crust_cache.clear();
    g_atm_cache.clear();
    for (auto& sl : g_slots) {
        sl->batch.clear();
        if (sl->ctx) dstar_ctx_destroy(sl->ctx);
    }
    g_slots.clear();
}

int dstar_alloc_NScool(void) {
    for (size_t i = 0; i < g_stars.size(); ++i)
        if (!g_stars[i] || !g_stars[i]->in_use) { g_stars[i].reset(new Star); g_stars[i]->in_use = true; return (int)i + 1; }
    if ((int)g_stars.size() >= max_handles) return -1;
    g_stars.emplace_back(new Star);
    g_stars.back()->in_use = true;
    return (int)g_stars.size();
}

int dstar_dealloc_NScool(int id) {
    Star* s = get(id);
    if (!s) return -1;
    g_stars[id - 1].reset();
    return 0;
}

int dstar_NScool_setup(int id, const char* inlist) {
    Star* s = get(id);
    if (!s) return -1;
    s->c = Controls();
    if (inlist && *inlist) {
        if (read_namelist(s->c, inlist, s->err) != 0) { g_err = s->err; return -1; }
    }
    s->setup = true;
    return 0;
}

int dstar_NScool_set_control(int id, const char* name, const char* value) {
    Star* s = get(id);
    if (!s || !name || !value) return -1;
    if (!s->setup) { s->c = Controls(); s->setup = true; }
    int rc = set_control(s->c, name, value);
    if (rc != 0) g_err = std::string("unknown or malformed control ") + name;
    return rc;
}

int dstar_NScool_set_hooks(int id, dstar_set_Qimp_fn qimp, dstar_sf_fn sf) {
    Star* s = get(id);
    if (!s) return -1;
    s->hook_qimp = qimp; s->hook_sf = sf;
    return 0;
}

namespace {
// owner[i] = first star of the batch whose listed per-zone arrays (and scalar tags) are byte-identical to star i's
void group_identical(const std::vector<Star*>& ss, const std::vector<const char*>& names, const std::vector<std::vector<double>>& extra,
                     const std::vector<int32_t>* seed, std::vector<int32_t>& owner) {
    const int n = (int)ss.size();
    std::map<std::string, int> first;
    owner.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        std::string key;
        auto add = [&](const void* p, size_t bytes) { key.append(reinterpret_cast<const char*>(p), bytes); };
        const int nz = ss[i]->nz;
        add(&nz, sizeof nz);
        if (seed) add(&(*seed)[i], sizeof(int32_t));
        for (const char* nm : names) { const auto& v = ss[i]->z[nm]; add(v.data(), v.size() * sizeof(double)); }
        if (!extra.empty()) add(extra[i].data(), extra[i].size() * sizeof(double));
        auto it = first.find(key);
        if (it == first.end()) { first.emplace(std::move(key), i); owner[i] = i; }
        else owner[i] = it->second;
    }
}
}  // namespace

int dstar_NScool_create_models(int n, const int32_t* ids) {
    if (!cur().ctx) { g_err = "NScool_init not called"; return -1; }
    std::vector<Star*> ss;
    for (int i = 0; i < n; ++i) { Star* s = get(ids[i]); if (!s || !s->setup) { g_err = "bad NScool id"; return -1; } ss.push_back(s); }
    cur().batch.clear();
    if (n > 1 && check_batch_controls(ss, true, false) != 0) return -1;
    // store_controls: epochs (NScool_ctrls_io.f:244-256), startup microphysics, zones
    for (Star* s : ss) {
        Controls& c = s->c;
        if (c.load_epochs) { if (load_epochs(c, s->epoch_boundaries, s->epoch_Mdots, s->err) != 0) { g_err = s->err; return -1; } }
        else {
            s->epoch_Mdots.resize(c.number_epochs); s->epoch_boundaries.resize(c.number_epochs + 1);
            for (int i = 0; i < c.number_epochs; ++i) s->epoch_Mdots[i] = c.basic_epoch_Mdots[i] * c.epoch_Mdot_scale;
            for (int i = 0; i <= c.number_epochs; ++i) s->epoch_boundaries[i] = c.basic_epoch_boundaries[i] * c.epoch_time_scale * k::julian_day;
        }
        s->number_epochs = c.number_epochs;
        if (ensure_gaps(c) != 0) return -1;
        if (build_crust_table(c, s->crust) != 0) return -1;
        if (setup_zones(*s) != 0) { g_err = s->err; return -1; }
        s->nisos = (int)s->crust->net.Z.size();
        s->ncharged = s->crust->net.ncharged();
    }
    if (build_bc09_tables(ss) != 0) return -1;
    // all stars of one batch must share the nuclide network
    const Network& net = ss[0]->crust->net;
    const int nch = ss[0]->ncharged;
    for (Star* s : ss) if (s->crust->net.names != net.names) { g_err = "stars of one batch must share the crust network"; return -1; }
    // do_setup_crust_composition (create_model.f:245-287): lookup on the device, per distinct crust table
    std::vector<int32_t> zoff(n + 1, 0);
    for (int i = 0; i < n; ++i) zoff[i + 1] = zoff[i] + ss[i]->nz;
    const int total = zoff[n];
    std::vector<double> ionic_bar((size_t)11 * total), Yion_bar((size_t)nch * total);
    {
        std::map<CrustTable*, std::vector<int>> groups;
        for (int i = 0; i < n; ++i) groups[ss[i]->crust.get()].push_back(i);
        for (auto& g : groups) {
            std::vector<double> lgP;
            for (int i : g.second) for (double p : ss[i]->z["P_bar"]) lgP.push_back(std::log10(p));
            std::vector<double> io((size_t)11 * lgP.size()), yi((size_t)nch * lgP.size());
            int32_t nc2 = 0;
            CrustTable* ct = g.first;
            int rc = dstar_crust_composition_info(cur().ctx, ct->nv, (int)ct->net.Z.size(), ct->lgP.data(), ct->Ycoef.data(),
                                                  ct->net.Z.data(), ct->net.A.data(), (int)lgP.size(), lgP.data(), io.data(),
                                                  yi.data(), &nc2);
            if (rc != 0) { g_err = dstar_last_error(); return rc; }
            size_t off = 0;
            for (int i : g.second) {
                std::memcpy(&ionic_bar[(size_t)11 * zoff[i]], &io[11 * off], sizeof(double) * 11 * ss[i]->nz);
                std::memcpy(&Yion_bar[(size_t)nch * zoff[i]], &yi[nch * off], sizeof(double) * nch * ss[i]->nz);
                off += ss[i]->nz;
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        Star* s = ss[i];
        auto& ib = s->z["ionic_bar"]; ib.assign(&ionic_bar[(size_t)11 * zoff[i]], &ionic_bar[(size_t)11 * zoff[i]] + 11 * s->nz);
        auto& yb = s->z["Yion_bar"]; yb.assign(&Yion_bar[(size_t)nch * zoff[i]], &Yion_bar[(size_t)nch * zoff[i]] + (size_t)nch * s->nz);
        for (int k2 = 0; k2 < s->nz; ++k2) s->z["Xneut_bar"][k2] = ib[11 * k2 + 9];
        if (s->c.fix_Qimp) for (int k2 = 0; k2 < s->nz; ++k2) ib[11 * k2 + 10] = s->c.Qimp;
        if (s->c.use_other_set_Qimp && s->hook_qimp) { int ierr = 0; s->hook_qimp(ids[i], &ierr); if (ierr) { g_err = "other_set_Qimp failed"; return ierr; } }
        s->z["Yion"] = yb; s->z["ionic"] = s->z["ionic_bar"]; s->z["Xneut"] = s->z["Xneut_bar"];
    }
    // do_setup_crust_transport on the device
    std::vector<double> rho, rho_bar, P, P_bar, dm, ionic, ionicb, Yion, Yionb, Zion, Aion;
    concat(ss, "rho", rho); concat(ss, "rho_bar", rho_bar); concat(ss, "P", P); concat(ss, "P_bar", P_bar); concat(ss, "dm", dm);
    concat(ss, "ionic", ionic); concat(ss, "ionic_bar", ionicb); concat(ss, "Yion", Yion); concat(ss, "Yion_bar", Yionb);
    for (size_t i = 0; i < net.Z.size(); ++i) if (net.Z[i] > 0) { Zion.push_back(net.Z[i]); Aion.push_back(net.A[i]); }
    // hook-evaluated critical temperatures for the cells (create_model.f:352-356)
    bool any_hook = false;
    for (Star* s : ss) any_hook |= (s->c.use_other_sf_critical_temperatures && s->hook_sf);
    auto eval_Tc = [&](bool face, const std::vector<double>& rb1, std::vector<double>& Tc) {
        Tc.assign((size_t)3 * total, 0.0);
        for (int i = 0; i < n; ++i) {
            Star* s = ss[i];
            const auto& io = s->z[face ? "ionic_bar" : "ionic"];
            for (int k2 = 0; k2 < s->nz; ++k2) {
                double r = face ? (k2 == 0 ? rb1[i] : s->z["rho_bar"][k2]) : s->z["rho"][k2];
                double Yn = io[11 * k2 + 9];
                double chi = k::onethird * k::fourpi * std::pow((double)1.13f * k::fm_to_cm, 3) * (r * (1.0 - Yn) / k::amu);
                double nn = r * Yn / k::amu / (1.0 - chi);
                double kn = std::pow(k::threepisquare * nn, k::onethird) / k::cm_to_fm;
                double* t = &Tc[(size_t)3 * (zoff[i] + k2)];
                if (s->c.use_other_sf_critical_temperatures && s->hook_sf) s->hook_sf(ids[i], 0.0, kn, t);
                else { t[0] = host_sf_Tc(0, 0.0); t[1] = host_sf_Tc(1, kn); t[2] = host_sf_Tc(2, kn); }
            }
        }
    };
    std::vector<double> Tc_cell, Tc_face, rb1(n, 0.0);
    if (any_hook) eval_Tc(false, rb1, Tc_cell);
    dstar_batch* b = nullptr;
    int rc = dstar_batch_create(cur().ctx, n, zoff.data(), nch, &b);
    if (rc != 0) { g_err = dstar_last_error(); return rc; }
    rc = dstar_batch_upload_structure(b, rho.data(), rho_bar.data(), P.data(), P_bar.data(), dm.data(), ionic.data(),
                                      ionicb.data(), Yion.data(), Yionb.data(), Zion.data(), Aion.data(),
                                      any_hook ? Tc_cell.data() : nullptr, nullptr);
    dstar_micro_ctrl mc = micro_ctrl_of(ss[0]->c);
    // table sharing (control share_identical_tables): a star's cell tables are a function of these inputs only
    // (create_model.f:344-403: rho, composition, T_c; P, P_bar, dm enter through the rho_bar(1) fix-up and shell Urca)
    bool share = n > 1;
    for (Star* s : ss) share = share && s->c.share_identical_tables;
    std::vector<int32_t> cell_owner, face_owner;
    cur().table_sets[0] = cur().table_sets[1] = n;
    if (share && rc == 0) {
        // the impurity parameter Q (11th moment) is carried by the cell composition too (create_model.f:283-286 copies
        // ionic_bar) but no cell quantity reads it -- EOS and crust neutrino emissivities take Z, A, Ye, Yn and the
        // charge moments -- so it stays out of the key (test_shared_tables_are_bit_identical holds this to the bit)
        std::vector<std::vector<double>> extra(n);
        for (int i = 0; i < n; ++i) {
            extra[i] = ss[i]->z["ionic"];
            for (int k2 = 0; k2 < ss[i]->nz; ++k2) extra[i][11 * k2 + 10] = 0.0;
            if (any_hook) extra[i].insert(extra[i].end(), &Tc_cell[(size_t)3 * zoff[i]], &Tc_cell[(size_t)3 * zoff[i]] + 3 * ss[i]->nz);
        }
        group_identical(ss, {"rho", "P", "P_bar", "dm", "Yion"}, extra, nullptr, cell_owner);
        rc = dstar_batch_set_table_sharing(b, 1, cell_owner.data());
        cur().table_sets[0] = 0;
        for (int i = 0; i < n; ++i) cur().table_sets[0] += cell_owner[i] == i;
    }
    if (rc == 0) rc = dstar_batch_micro_tables(b, &mc, 1);
    if (rc == 0) { cur().ms[0] = dstar_batch_last_kernel_ms(b, 0); rc = dstar_batch_get_rho_bar1(b, rb1.data()); }
    if (rc == 0 && any_hook) { eval_Tc(true, rb1, Tc_face); rc = dstar_batch_set_Tc_face(b, Tc_face.data()); }
    if (any_hook)
        for (int i = 0; i < n; ++i) {  // keep the hook-evaluated critical temperatures with the star
            ss[i]->z["Tc_cell"].assign(&Tc_cell[(size_t)3 * zoff[i]], &Tc_cell[(size_t)3 * zoff[i]] + 3 * ss[i]->nz);
            if (!Tc_face.empty())
                ss[i]->z["Tc_face"].assign(&Tc_face[(size_t)3 * zoff[i]], &Tc_face[(size_t)3 * zoff[i]] + 3 * ss[i]->nz);
        }
    if (share && rc == 0) {
        // face tables (create_model.f:406-441): rho_bar (element 1 follows from the cell inputs: the cell owner stands
        // for it), the face composition with the impurity parameter, the face T_c
        std::vector<std::vector<double>> extra;
        if (any_hook) for (int i = 0; i < n; ++i) extra.emplace_back(&Tc_face[(size_t)3 * zoff[i]], &Tc_face[(size_t)3 * zoff[i]] + 3 * ss[i]->nz);
        std::vector<double> keep(n);
        for (int i = 0; i < n; ++i) { keep[i] = ss[i]->z["rho_bar"][0]; ss[i]->z["rho_bar"][0] = 0.0; }
        group_identical(ss, {"rho_bar", "ionic_bar", "Yion_bar"}, extra, &cell_owner, face_owner);
        for (int i = 0; i < n; ++i) ss[i]->z["rho_bar"][0] = keep[i];
        rc = dstar_batch_set_table_sharing(b, 2, face_owner.data());
        cur().table_sets[1] = 0;
        for (int i = 0; i < n; ++i) cur().table_sets[1] += face_owner[i] == i;
    }
    if (rc == 0) rc = dstar_batch_micro_tables(b, &mc, 2);
    if (rc == 0) cur().ms[1] = dstar_batch_last_kernel_ms(b, 1);
    std::vector<double> tlnT(DSTAR_NTAB), tCp((size_t)4 * DSTAR_NTAB * total), tGa(tCp.size()), tEn(tCp.size()), tK(tCp.size());
    std::vector<int32_t> status(n);
    if (rc == 0) rc = dstar_batch_download_tables(b, tlnT.data(), tCp.data(), tGa.data(), tEn.data(), tK.data(), rb1.data(), status.data());
    if (rc != 0) { g_err = dstar_last_error(); dstar_batch_destroy(b); return rc; }
    for (int i = 0; i < n; ++i) {
        Star* s = ss[i];
        const size_t o = (size_t)4 * DSTAR_NTAB * zoff[i], len = (size_t)4 * DSTAR_NTAB * s->nz;
        s->z["tab_lnT"] = tlnT;
        s->z["tab_lnCp"].assign(&tCp[o], &tCp[o] + len); s->z["tab_lnGamma"].assign(&tGa[o], &tGa[o] + len);
        s->z["tab_lnEnu"].assign(&tEn[o], &tEn[o] + len); s->z["tab_lnK"].assign(&tK[o], &tK[o] + len);
        s->z["rho_bar"][0] = rb1[i];
        s->created = true; s->tsec = 0; s->dt = 0; s->model = 0;
        if (status[i] != 0) { g_err = "error in helmeos2 while tabulating coefficients"; dstar_batch_destroy(b); return -1; }
    }
    cur().batch.b = b;
    cur().batch.ids.assign(ids, ids + n);
    return 0;
}

int dstar_NScool_create_model(int id) { int32_t i = id; return dstar_NScool_create_models(1, &i); }

int dstar_NScool_evolve_models(int n, const int32_t* ids) {
    if (!cur().ctx) { g_err = "NScool_init not called"; return -1; }
    std::vector<Star*> ss;
    for (int i = 0; i < n; ++i) { Star* s = get(ids[i]); if (!s || !s->created) { g_err = "model not created"; return -1; } ss.push_back(s); }
    const int ne = ss[0]->number_epochs;
    for (Star* s : ss) if (s->number_epochs != ne) { g_err = "stars of one batch must have the same number of epochs"; return -1; }
    if (n > 1 && check_batch_controls(ss, false, true) != 0) return -1;
    std::vector<int32_t> zoff(n + 1, 0);
    for (int i = 0; i < n; ++i) zoff[i + 1] = zoff[i] + ss[i]->nz;
    const int total = zoff[n];
    dstar_batch* b = nullptr;
    int rc = 0;
    bool reuse = cur().batch.b && cur().batch.ids == std::vector<int>(ids, ids + n);
    if (reuse) b = cur().batch.b;
    else {
        cur().batch.clear();
        const Network& net = ss[0]->crust->net;
        std::vector<double> rho, rho_bar, P, P_bar, dm, ionic, ionicb, Yion, Yionb, Zion, Aion, tCp, tGa, tEn, tK;
        concat(ss, "rho", rho); concat(ss, "rho_bar", rho_bar); concat(ss, "P", P); concat(ss, "P_bar", P_bar); concat(ss, "dm", dm);
        concat(ss, "ionic", ionic); concat(ss, "ionic_bar", ionicb); concat(ss, "Yion", Yion); concat(ss, "Yion_bar", Yionb);
        for (size_t i = 0; i < net.Z.size(); ++i) if (net.Z[i] > 0) { Zion.push_back(net.Z[i]); Aion.push_back(net.A[i]); }
        concat(ss, "tab_lnCp", tCp); concat(ss, "tab_lnGamma", tGa); concat(ss, "tab_lnEnu", tEn); concat(ss, "tab_lnK", tK);
        rc = dstar_batch_create(cur().ctx, n, zoff.data(), ss[0]->ncharged, &b);
        if (rc == 0) rc = dstar_batch_upload_structure(b, rho.data(), rho_bar.data(), P.data(), P_bar.data(), dm.data(), ionic.data(),
                                                       ionicb.data(), Yion.data(), Yionb.data(), Zion.data(), Aion.data(), nullptr, nullptr);
        if (rc == 0) rc = dstar_batch_upload_tables(b, nullptr, tCp.data(), tGa.data(), tEn.data(), tK.data());
        if (rc != 0) { g_err = dstar_last_error(); if (b) dstar_batch_destroy(b); return rc; }
        cur().batch.b = b; cur().batch.ids.assign(ids, ids + n);
    }
    std::vector<double> dm_bar, area, ePhi, e2Phi, ePhi_bar, e2Phi_bar, lnT0;
    concat(ss, "dm_bar", dm_bar); concat(ss, "area", area); concat(ss, "ePhi", ePhi); concat(ss, "e2Phi", e2Phi);
    concat(ss, "ePhi_bar", ePhi_bar); concat(ss, "e2Phi_bar", e2Phi_bar); concat(ss, "lnT", lnT0);
    const int nv = ss[0]->atm.nv;
    std::vector<double> aTb, aTe, aFl, heat((size_t)9 * n), T_atm(n), eb((size_t)(ne + 1) * n), em((size_t)ne * n);
    std::vector<int32_t> aidx(n);
    for (int i = 0; i < n; ++i) {
        Star* s = ss[i];
        aidx[i] = i;
        aTb.insert(aTb.end(), s->atm.lgTb.begin(), s->atm.lgTb.end());
        aTe.insert(aTe.end(), s->atm.lgTeff.begin(), s->atm.lgTeff.end());
        aFl.insert(aFl.end(), s->atm.lgflux.begin(), s->atm.lgflux.end());
        const Controls& c = s->c;
        double h[9] = {c.lgP_min_heating_outer, c.lgP_max_heating_outer, c.Q_heating_outer, c.lgP_min_heating_inner,
                       c.lgP_max_heating_inner, c.Q_heating_inner, c.lgP_min_heating_shallow, c.lgP_max_heating_shallow,
                       c.Q_heating_shallow};
        std::memcpy(&heat[(size_t)9 * i], h, sizeof h);
        T_atm[i] = c.atmosphere_temperature_when_accreting;
        for (int e = 0; e <= ne; ++e) eb[(size_t)e * n + i] = s->epoch_boundaries[e];
        for (int e = 0; e < ne; ++e) em[(size_t)e * n + i] = s->epoch_Mdots[e];
    }
    rc = dstar_batch_upload_evolve(b, dm_bar.data(), area.data(), ePhi.data(), e2Phi.data(), ePhi_bar.data(), e2Phi_bar.data(),
                                   n, nv, aTb.data(), aTe.data(), aFl.data(), aidx.data(), heat.data(), T_atm.data(),
                                   lnT0.data(), ne, eb.data(), em.data(), nullptr);
    dstar_evolve_ctrl ec = evolve_ctrl_of(ss[0]->c);
    int cap = ss[0]->c.trace_capacity;
    if (cap < 0) cap = (n <= 64) ? 2048 : 0;
    int pcap = ss[0]->c.profile_capacity;
    if (pcap < 0) pcap = (n <= 64) ? 512 : 0;
    const int pint = std::max(1, ss[0]->c.write_interval_for_profile);
    if (rc == 0) rc = dstar_batch_set_profile_capture(b, pcap > 0 ? pint : 0, pcap);
    if (rc == 0) rc = dstar_batch_set_evolve_mapping(b, g_evolve_mapping);
    if (rc == 0) rc = dstar_batch_evolve(b, &ec, cap);
    if (rc == 0) cur().ms[2] = dstar_batch_last_kernel_ms(b, 2);
    std::vector<double> tm((size_t)ne * n), Tm(tm.size()), Qm(tm.size()), lnT(total), Teff(n), Lsurf(n), dt(n), tsec(n);
    std::vector<int32_t> idid((size_t)ne * n), cnt((size_t)7 * n), model(n);
    if (rc == 0) rc = dstar_batch_download_results(b, tm.data(), Tm.data(), Qm.data(), idid.data(), cnt.data(), lnT.data(),
                                                   Teff.data(), Lsurf.data(), dt.data(), tsec.data(), model.data());
    std::vector<int32_t> tc, tmodel;
    std::vector<double> ttsec, tdt, tTeff, tLs, tLnu, tLnuc, tMdot, text;
    if (rc == 0 && cap > 0) {
        size_t tn = (size_t)cap * n;
        tc.resize(n); tmodel.resize(tn); ttsec.resize(tn); tdt.resize(tn); tTeff.resize(tn); tLs.resize(tn); tLnu.resize(tn);
        tLnuc.resize(tn); tMdot.resize(tn);
        rc = dstar_batch_download_trace(b, tc.data(), tmodel.data(), ttsec.data(), tdt.data(), tTeff.data(), tLs.data(),
                                        tLnu.data(), tLnuc.data(), tMdot.data());
        text.resize(4 * tn);
        if (rc == 0) rc = dstar_batch_download_trace_extrema(b, text.data());
    }
    std::vector<int32_t> pc, pmodel;
    std::vector<double> ptsec, pMdot, plnT;
    if (rc == 0 && pcap > 0) {
        pc.resize(n); pmodel.resize((size_t)pcap * n); ptsec.resize(pmodel.size()); pMdot.resize(pmodel.size());
        plnT.resize((size_t)pcap * total);
        rc = dstar_batch_download_profiles(b, pc.data(), pmodel.data(), ptsec.data(), pMdot.data(), plnT.data());
    }
    if (rc != 0) { g_err = dstar_last_error(); return rc; }
    int worst = 0;
    for (int i = 0; i < n; ++i) {
        Star* s = ss[i];
        s->t_monitor.resize(ne); s->Teff_monitor.resize(ne); s->Qb_monitor.resize(ne); s->idid.resize(ne);
        for (int e = 0; e < ne; ++e) {
            s->t_monitor[e] = tm[(size_t)e * n + i]; s->Teff_monitor[e] = Tm[(size_t)e * n + i];
            s->Qb_monitor[e] = Qm[(size_t)e * n + i]; s->idid[e] = idid[(size_t)e * n + i];
            if (s->idid[e] < 0) worst = s->idid[e];
        }
        std::memcpy(s->counters, &cnt[(size_t)7 * i], sizeof(int) * 7);
        std::copy(&lnT[zoff[i]], &lnT[zoff[i]] + s->nz, s->z["lnT"].begin());
        for (int k2 = 0; k2 < s->nz; ++k2) s->z["T"][k2] = std::exp(s->z["lnT"][k2]);
        s->Teff = Teff[i]; s->Lsurf = Lsurf[i]; s->dt = dt[i]; s->tsec = tsec[i]; s->model = model[i];
        s->h_model.clear(); s->h_tsec.clear(); s->h_Mdot.clear(); s->h_Teff.clear(); s->h_Lsurf.clear(); s->h_Lnu.clear(); s->h_Lnuc.clear();
        s->h_dt.clear(); s->h_lnTmax.clear(); s->h_PTmax.clear(); s->h_lnTmin.clear(); s->h_PTmin.clear(); s->h_total = 0;
        if (cap > 0) {
            s->h_total = tc[i];
            int rows = std::min(tc[i], cap);
            const size_t cs = (size_t)cap * n;
            for (int r = 0; r < rows; ++r) {
                size_t q = (size_t)r * n + i;
                s->h_model.push_back(tmodel[q]); s->h_tsec.push_back(ttsec[q]); s->h_Mdot.push_back(tMdot[q]);
                s->h_Teff.push_back(tTeff[q]); s->h_Lsurf.push_back(tLs[q]); s->h_Lnu.push_back(tLnu[q]); s->h_Lnuc.push_back(tLnuc[q]);
                s->h_dt.push_back(tdt[q]); s->h_lnTmax.push_back(text[q]); s->h_PTmax.push_back(text[cs + q]);
                s->h_lnTmin.push_back(text[2 * cs + q]); s->h_PTmin.push_back(text[3 * cs + q]);
            }
        }
        s->p_model.clear(); s->p_tsec.clear(); s->p_Mdot.clear(); s->p_lnT.clear(); s->p_total = 0;
        if (pcap > 0) {
            s->p_total = pc[i];
            const int rows = std::min(pc[i], pcap);
            for (int r = 0; r < rows; ++r) {
                const size_t q = (size_t)r * n + i;
                s->p_model.push_back(pmodel[q]); s->p_tsec.push_back(ptsec[q]); s->p_Mdot.push_back(pMdot[q]);
                const double* src = &plnT[(size_t)r * total + zoff[i]];
                s->p_lnT.insert(s->p_lnT.end(), src, src + s->nz);
            }
        }
        s->evolved = true;
    }
    return worst;
}

int dstar_NScool_evolve_model(int id) { int32_t i = id; return dstar_NScool_evolve_models(1, &i); }

// The sweep driver over the GPUs of one node: what N processes calling NScool_create_model / NScool_evolve_model do in
// the reference (NScool/public/NScool_lib.f:40-90, one star per call).  Stars are dealt to the devices in contiguous
// blocks (sizes differ by at most one, as dstar_b200/dist.py shard_range), one host thread per device runs
// create_models + evolve_models on its block; nothing is exchanged between the devices -- stars do not interact
// (NScool_def.f:160) -- and the results stay in the star handles.  Monitors are also returned concatenated.
int dstar_NScool_run_batch(int n_gpu, const int32_t* devices, int n, const int32_t* ids, double* t_monitor,
                           double* Teff_monitor, double* Qb_monitor, int32_t* star_status) {
    if (g_slots.empty()) { g_err = "NScool_init not called"; return -1; }
    if (n_gpu < 1 || n < 1 || !ids) { g_err = "run_batch: n_gpu >= 1, n >= 1 and an id list are required"; return -1; }
    if (t_slot) { g_err = "run_batch: not re-entrant"; return -1; }
    // devices == NULL: the device of NScool_init, then the other devices of the node in index order
    std::vector<int> devs;
    if (devices) devs.assign(devices, devices + n_gpu);
    else {
        devs.push_back(g_slots[0]->device);
        for (int d = 0; (int)devs.size() < n_gpu; ++d) if (d != g_slots[0]->device) devs.push_back(d);
    }
    std::vector<DevSlot*> slots(n_gpu, nullptr);
    for (int g = 0; g < n_gpu; ++g) {
        for (int h = 0; h < g; ++h) if (devs[h] == devs[g]) { g_err = "run_batch: a device is listed twice"; return -1; }
        int rc = open_slot(devs[g], &slots[g]);
        if (rc != 0) return rc;
    }
    std::vector<int> rcs(n_gpu, 0);
    std::vector<std::string> errs(n_gpu);
    std::vector<std::thread> th;
    const int base = n / n_gpu, rem = n % n_gpu;
    std::vector<int> lo(n_gpu + 1, 0);
    for (int g = 0; g < n_gpu; ++g) lo[g + 1] = lo[g] + base + (g < rem ? 1 : 0);
    for (int g = 0; g < n_gpu; ++g) {
        th.emplace_back([&, g] {
            t_slot = slots[g];
            const int cnt = lo[g + 1] - lo[g];
            int rc = 0;
            if (cnt > 0) {
                rc = dstar_NScool_create_models(cnt, ids + lo[g]);
                if (rc == 0) rc = dstar_NScool_evolve_models(cnt, ids + lo[g]);
                if (rc < 0) errs[g] = g_err;
            }
            rcs[g] = rc;
            t_slot = nullptr;
        });
    }
    for (auto& t : th) t.join();
    int worst = 0;
    for (int g = 0; g < n_gpu; ++g) {
        if (rcs[g] < 0 && worst >= 0) { worst = rcs[g]; g_err = "device " + std::to_string(slots[g]->device) + ": " + errs[g]; }
        else if (rcs[g] > worst && worst >= 0) worst = rcs[g];
    }
    // what the last batch call reports for the whole job: distinct table sets summed, kernel times the slowest device's
    int sets[2] = {0, 0};
    double ms[3] = {0, 0, 0};
    for (int g = 0; g < n_gpu; ++g) {
        if (lo[g + 1] == lo[g]) continue;
        for (int q = 0; q < 2; ++q) sets[q] += slots[g]->table_sets[q];
        for (int q = 0; q < 3; ++q) ms[q] = std::max(ms[q], slots[g]->ms[q]);
    }
    for (int q = 0; q < 2; ++q) g_slots[0]->table_sets[q] = sets[q];
    for (int q = 0; q < 3; ++q) g_slots[0]->ms[q] = ms[q];
    if (worst < 0) return worst;
    for (int i = 0; i < n; ++i) {
        Star* s = get(ids[i]);
        if (!s) continue;
        for (int e = 0; e < s->number_epochs; ++e) {   // (number_epochs, n): epoch-major, as the batch kernels keep them
            if (t_monitor) t_monitor[(size_t)e * n + i] = s->t_monitor[e];
            if (Teff_monitor) Teff_monitor[(size_t)e * n + i] = s->Teff_monitor[e];
            if (Qb_monitor) Qb_monitor[(size_t)e * n + i] = s->Qb_monitor[e];
        }
        if (star_status) { int w = 0; for (int v : s->idid) w = std::min(w, v); star_status[i] = w; }
    }
    return worst;
}

int dstar_NScool_table_sets(int32_t* cell_sets, int32_t* face_sets) {
    if (cell_sets) *cell_sets = cur().table_sets[0];
    if (face_sets) *face_sets = cur().table_sets[1];
    return 0;
}

int dstar_NScool_set_evolve_mapping(int mapping) {
    if (mapping < 0 || mapping > 2) { g_err = "evolve mapping: 0 automatic, 1 CTA per star, 2 warp per star"; return -1; }
    g_evolve_mapping = mapping;
    return 0;
}

int dstar_NScool_get_int(int id, const char* name, int32_t* v) {
    Star* s = get(id);
    if (!s || !name || !v) return -1;
    std::string nm = name;
    if (nm == "nz") *v = s->nz; else if (nm == "number_epochs") *v = s->number_epochs; else if (nm == "ncharged") *v = s->ncharged;
    else if (nm == "nisos") *v = s->nisos; else if (nm == "model") *v = s->model; else return -2;
    return 0;
}

int dstar_NScool_get_real(int id, const char* name, double* v) {
    Star* s = get(id);
    if (!s || !name || !v) return -1;
    std::string nm = name;
    const std::map<std::string, double> m = {{"grav", s->grav}, {"Mtotal", s->Mtotal}, {"Rtotal", s->Rtotal}, {"Teff", s->Teff},
        {"Lsurf", s->Lsurf}, {"Tcore", s->Tcore}, {"Psurf", s->Psurf}, {"Pcore", s->Pcore}, {"Mcore", s->Mcore}, {"Rcore", s->Rcore},
        {"ePhicore", s->ePhicore}, {"tsec", s->tsec}, {"dt", s->dt}};
    auto it = m.find(nm);
    if (it == m.end()) return -2;
    *v = it->second;
    return 0;
}

int dstar_NScool_get_zone_array(int id, const char* name, double* out, int capacity) {
    Star* s = get(id);
    if (!s || !name) return -1;
    std::string nm = name;
    std::vector<double> tmp;
    const std::vector<double>* src = nullptr;
    if (nm == "epoch_boundaries") src = &s->epoch_boundaries;
    else if (nm == "epoch_Mdots") src = &s->epoch_Mdots;
    else if (nm == "atm_lgTb") src = &s->atm.lgTb;
    else if (nm == "atm_lgTeff") src = &s->atm.lgTeff;
    else if (nm == "atm_lgflux") src = &s->atm.lgflux;
    else if (nm == "Zion" || nm == "Aion") {
        const Network& net = s->crust->net;
        for (size_t i = 0; i < net.Z.size(); ++i) if (net.Z[i] > 0) tmp.push_back(nm == "Zion" ? net.Z[i] : net.A[i]);
        src = &tmp;
    } else if (nm.rfind("crust_", 0) == 0 && s->crust) {   // the star's crust table (dStar_crust_def.f:10-24)
        const CrustTable& ct = *s->crust;
        if (nm == "crust_lgP") src = &ct.lgP;
        else if (nm == "crust_lgRho") src = &ct.lgRho;      // (4, nv) monotone-cubic coefficients
        else if (nm == "crust_lgEps") src = &ct.lgEps;
        else if (nm == "crust_ionic") src = &ct.ionic;
        else if (nm == "crust_Yion") src = &ct.Yion;
        else if (nm == "crust_mass_excess") src = &ct.mass_excess;
        else if (nm == "crust_Tref") { tmp.push_back(ct.T); src = &tmp; }
        else return -2;
    } else if (nm == "Qimp_bar") {
        auto& ib = s->z["ionic_bar"];
        for (int k2 = 0; k2 < s->nz; ++k2) tmp.push_back(ib[11 * k2 + 10]);
        src = &tmp;
    } else {
        auto it = s->z.find(nm);
        if (it == s->z.end()) return -2;
        src = &it->second;
    }
    if (out) for (int i = 0; i < capacity && i < (int)src->size(); ++i) out[i] = (*src)[i];
    return (int)src->size();
}

int dstar_NScool_set_zone_array(int id, const char* name, const double* in, int count) {
    Star* s = get(id);
    if (!s || !name || !in) return -1;
    std::string nm = name;
    if (nm == "Qimp_bar") {
        auto& ib = s->z["ionic_bar"];
        for (int k2 = 0; k2 < s->nz && k2 < count; ++k2) ib[11 * k2 + 10] = in[k2];
        return 0;
    }
    auto it = s->z.find(nm);
    if (it == s->z.end() || (int)it->second.size() != count) return -2;
    std::copy(in, in + count, it->second.begin());
    if (cur().batch.b && std::find(cur().batch.ids.begin(), cur().batch.ids.end(), id) != cur().batch.ids.end()) cur().batch.clear();
    return 0;
}

int dstar_NScool_get_monitors(int id, double* t, double* Teff, double* Qb) {
    Star* s = get(id);
    if (!s || !s->evolved) return -1;
    for (int e = 0; e < s->number_epochs; ++e) {
        if (t) t[e] = s->t_monitor[e];
        if (Teff) Teff[e] = s->Teff_monitor[e];
        if (Qb) Qb[e] = s->Qb_monitor[e];
    }
    return s->number_epochs;
}

int dstar_NScool_get_counters(int id, int32_t c[7]) {
    Star* s = get(id);
    if (!s || !s->evolved) return -1;
    for (int i = 0; i < 7; ++i) c[i] = s->counters[i];
    return 0;
}

static double safe_log10(double x) { return std::log10(std::max(1.0e-99, x)); }  // MESA math_lib safe_log10

int dstar_NScool_get_history(int id, int capacity, int32_t* model, double* time_d, double* lg_Mdot, double* lg_Teff,
                             double* lg_Lsurf, double* lg_Lnu, double* lg_Lnuc) {
    Star* s = get(id);
    if (!s || !s->evolved) return -1;
    int rows = (int)s->h_model.size();
    for (int r = 0; r < rows && r < capacity; ++r) {
        if (model) model[r] = s->h_model[r];
        if (time_d) time_d[r] = s->h_tsec[r] / k::julian_day;
        if (lg_Mdot) lg_Mdot[r] = safe_log10(s->h_Mdot[r]);
        if (lg_Teff) lg_Teff[r] = std::log10(s->h_Teff[r]);
        if (lg_Lsurf) lg_Lsurf[r] = std::log10(s->h_Lsurf[r]);
        if (lg_Lnu) lg_Lnu[r] = std::log10(s->h_Lnu[r]);
        if (lg_Lnuc) lg_Lnuc[r] = safe_log10(s->h_Lnuc[r]);
    }
    return rows;
}

// NScool_history.f:25-89: header block, then one row per model (write_interval_for_history)
int dstar_NScool_write_history(int id, const char* path) {
    Star* s = get(id);
    if (!s || !s->evolved || !path) return -1;
    if (s->h_total == 0) { g_err = "no history rows were kept (trace_capacity = 0 for this batch: set the control trace_capacity)"; return -1; }
    FILE* fp = std::fopen(path, "w");
    if (!fp) return -1;
    std::fprintf(fp, "dstar-b200/history\n\n");
    for (int i = 1; i <= 6; ++i) std::fprintf(fp, "%15d", i);
    std::fprintf(fp, "\n%15s%15s%15s%15s%15s%15s\n", "gravity", "total_mass", "total_radius", "core_mass", "core_radius", "core_temp");
    std::fprintf(fp, "%15.6E%15.6f%15.6f%15.6f%15.6f%15.6E\n\n\n", s->grav, s->Mtotal, s->Rtotal, s->Mcore, s->Rcore, s->Tcore);
    for (int i = 1; i <= 7; ++i) std::fprintf(fp, "%15d", i);
    std::fprintf(fp, "\n%15s%15s%15s%15s%15s%15s%15s\n", "model", "time", "lg_Mdot", "lg_Teff", "lg_Lsurf", "lg_Lnu", "lg_Lnuc");
    const int every = std::max(1, s->c.write_interval_for_history);
    for (size_t r = 0; r < s->h_model.size(); ++r) {
        if (s->h_model[r] % every != 0) continue;
        std::fprintf(fp, "%15d%15.6E%15.6f%15.6f%15.6f%15.6f%15.6f\n", s->h_model[r], s->h_tsec[r] / k::julian_day,
                     safe_log10(s->h_Mdot[r]), std::log10(s->h_Teff[r]), std::log10(s->h_Lsurf[r]), std::log10(s->h_Lnu[r]),
                     safe_log10(s->h_Lnuc[r]));
    }
    std::fclose(fp);
    if (s->h_total > (int)s->h_model.size()) {   // the file is complete up to the capacity only: tell the caller
        g_err = "history truncated: " + std::to_string(s->h_total) + " rows produced, " + std::to_string(s->h_model.size()) +
                " kept (raise the control trace_capacity)";
        return 1;
    }
    return 0;
}

// NScool_terminal.f:17-57 as evaluate_timestep drives it (NScool_evolve.f:206-217): one line per model with
// mod(model, write_interval_for_terminal) == 0, the column titles before every write_interval_for_terminal_header-th
// line of an epoch (the counter restarts with each epoch, :109).  path NULL or "-": standard output.
int dstar_NScool_write_terminal(int id, const char* path) {
    Star* s = get(id);
    if (!s || !s->evolved) return -1;
    if (s->h_total == 0) { g_err = "no history rows were kept (trace_capacity = 0 for this batch: set the control trace_capacity)"; return -1; }
    const bool to_stdout = !path || std::strcmp(path, "-") == 0;
    FILE* fp = to_stdout ? stdout : std::fopen(path, "w");
    if (!fp) return -1;
    static const char* cols[11] = {"model", "time", "dt", "lg(Lsurf)", "lg(Teff)", "lg(Lnu)", "lg(Lnuc)", "max(lgT)", "lg(P[Tmax])",
                                   "min(lgT)", "lg(P[Tmin])"};
    const int every = std::max(1, s->c.write_interval_for_terminal), hdr = std::max(1, s->c.write_interval_for_terminal_header);
    int epoch = 0, writes = 0;
    for (size_t r = 0; r < s->h_model.size(); ++r) {
        // a new epoch: its first row repeats the boundary time with dt = 0, or (first output suppressed) lies past it
        while (epoch + 1 < s->number_epochs &&
               ((r > 0 && s->h_dt[r] == 0.0 && s->h_tsec[r] >= s->epoch_boundaries[epoch + 1]) || s->h_tsec[r] > s->epoch_boundaries[epoch + 1])) {
            ++epoch; writes = 0;
        }
        if (s->h_model[r] % every != 0) continue;
        if (writes % hdr == 0) {
            std::fprintf(fp, "\n");
            for (int i = 0; i < 11; ++i) std::fprintf(fp, "%15s", cols[i]);
            std::fprintf(fp, "\n\n");
        }
        std::fprintf(fp, "%15d%15.6E%15.6E%15.6f%15.6f%15.6f%15.6f%15.6f%15.6f%15.6f%15.6f\n", s->h_model[r], s->h_tsec[r], s->h_dt[r],
                     std::log10(s->h_Lsurf[r]), std::log10(s->h_Teff[r]), std::log10(s->h_Lnu[r]), safe_log10(s->h_Lnuc[r]),
                     s->h_lnTmax[r] / k::ln10, std::log10(s->h_PTmax[r]), s->h_lnTmin[r] / k::ln10, std::log10(s->h_PTmin[r]));
        ++writes;
    }
    if (!to_stdout) std::fclose(fp);
    else std::fflush(stdout);
    if (s->h_total > (int)s->h_model.size()) { g_err = "terminal log truncated (raise the control trace_capacity)"; return 1; }
    return 0;
}

int dstar_NScool_get_terminal(int id, int capacity, int32_t* model, double* tsec, double* dt, double* lgTmax,
                              double* lgP_Tmax, double* lgTmin, double* lgP_Tmin) {
    Star* s = get(id);
    if (!s || !s->evolved) return -1;
    const int rows = (int)s->h_model.size();
    for (int r = 0; r < rows && r < capacity; ++r) {
        if (model) model[r] = s->h_model[r];
        if (tsec) tsec[r] = s->h_tsec[r];
        if (dt) dt[r] = s->h_dt[r];
        if (lgTmax) lgTmax[r] = s->h_lnTmax[r] / k::ln10;
        if (lgP_Tmax) lgP_Tmax[r] = std::log10(s->h_PTmax[r]);
        if (lgTmin) lgTmin[r] = s->h_lnTmin[r] / k::ln10;
        if (lgP_Tmin) lgP_Tmin[r] = std::log10(s->h_PTmin[r]);
    }
    return rows;
}

// ---- profiles (NScool_profile.f): the device keeps ln T per zone at the profile models; every other column is
// rebuilt here from the star's structure and coefficient tables exactly as evaluate_timestep refreshes the state
// (interpolate_temps :452-459, get_coefficients :461-524, evaluate_luminosity :433-450, get_nuclear_heating :526-568)
namespace {
void profile_columns(Star* s, int ip, std::vector<double>& col /* (20, nz) row-major by zone: col[20*iz + c] */, double& Lnu, double& Lnuc) {
    const int nz = s->nz, nt = 128;
    auto& Z = s->z;
    const double* lnT = &s->p_lnT[(size_t)ip * nz];
    const double Mdot = s->p_Mdot[ip];
    const std::vector<double>& tab = Z["tab_lnT"];
    std::vector<double> T(nz), K(nz), Cp(nz), Gam(nz), enu(nz), enuc(nz, 0.0), L(nz);
    for (int i = 0; i < nz; ++i) T[i] = std::exp(lnT[i]);
    for (int i = 0; i < nz; ++i) {
        const double Tbar = (i == 0) ? T[0] : 0.5 * (T[i - 1] * Z["dm"][i] + T[i] * Z["dm"][i - 1]) / Z["dm_bar"][i];
        const double lnTbar = std::log(Tbar);
        K[i] = std::exp(interp_value(tab.data(), nt, &Z["tab_lnK"][(size_t)4 * nt * i], lnTbar));
        Cp[i] = std::exp(interp_value(tab.data(), nt, &Z["tab_lnCp"][(size_t)4 * nt * i], lnT[i]));
        Gam[i] = std::exp(interp_value(tab.data(), nt, &Z["tab_lnGamma"][(size_t)4 * nt * i], lnT[i]));
        enu[i] = std::exp(interp_value(tab.data(), nt, &Z["tab_lnEnu"][(size_t)4 * nt * i], lnT[i]));
        if (i == 0) {
            double lgTb = lnTbar / k::ln10;
            lgTb = std::min(std::max(lgTb, s->atm.lgTb.front()), s->atm.lgTb.back());
            L[0] = Z["area"][0] * std::pow(10.0, interp_value(s->atm.lgTb.data(), s->atm.nv, s->atm.lgflux.data(), lgTb));
        } else {
            L[i] = -(Z["area"][i] * Z["area"][i]) * Z["rho_bar"][i] * K[i] * (Z["ePhi"][i - 1] * T[i - 1] - Z["ePhi"][i] * T[i]) /
                   Z["ePhi_bar"][i] / Z["dm_bar"][i];
        }
    }
    const Controls& c = s->c;
    const double hw[9] = {c.lgP_min_heating_outer, c.lgP_max_heating_outer, c.Q_heating_outer, c.lgP_min_heating_inner,
                          c.lgP_max_heating_inner, c.Q_heating_inner, c.lgP_min_heating_shallow, c.lgP_max_heating_shallow,
                          c.Q_heating_shallow};
    for (int w = 0; w < (c.turn_on_extra_heating ? 3 : 2); ++w) {
        const double Ptop = std::pow(10.0, hw[3 * w]), Pbot = std::pow(10.0, hw[3 * w + 1]);
        double M = 0;
        for (int i = 0; i < nz; ++i) if (Z["P"][i] >= Ptop && Z["P"][i] <= Pbot) M += Z["dm"][i];
        for (int i = 0; i < nz; ++i) if (Z["P"][i] >= Ptop && Z["P"][i] <= Pbot) enuc[i] = Mdot * hw[3 * w + 2] * k::mev_to_ergs * k::avogadro / M;
    }
    Lnu = 0; Lnuc = 0;
    for (int i = 0; i < nz; ++i) { Lnu += enu[i] * Z["dm"][i]; Lnuc += enuc[i] * Z["dm"][i]; }
    col.assign((size_t)20 * nz, 0.0);
    for (int i = 0; i < nz; ++i) {
        double* q = &col[(size_t)20 * i];
        const double* io = &Z["ionic"][(size_t)11 * i];
        const double grav = k::fourpi * k::Gnewton * k::Msun * s->Mcore * Z["eLambda"][i] / Z["area"][i];
        q[0] = i + 1; q[1] = Z["m"][i]; q[2] = Z["dm"][i]; q[3] = Z["area"][i]; q[4] = grav; q[5] = Z["eLambda"][i];
        q[6] = Z["ePhi"][i]; q[7] = T[i]; q[8] = L[i]; q[9] = Z["P_bar"][i]; q[10] = Z["rho"][i]; q[11] = io[1]; q[12] = io[0];
        q[13] = Z["Xneut"][i]; q[14] = io[10]; q[15] = Gam[i]; q[16] = Cp[i]; q[17] = enu[i]; q[18] = enuc[i]; q[19] = K[i];
    }
}
}  // namespace

int dstar_NScool_profile_count(int id, int32_t* kept, int32_t* produced) {
    Star* s = get(id);
    if (!s || !s->evolved) return -1;
    if (kept) *kept = (int)s->p_model.size();
    if (produced) *produced = s->p_total;
    return 0;
}

int dstar_NScool_get_profile(int id, int index, int32_t* model, double* header /* 8 */, double* columns /* 20*nz */) {
    Star* s = get(id);
    if (!s || !s->evolved || index < 0 || index >= (int)s->p_model.size()) return -1;
    std::vector<double> col;
    double Lnu, Lnuc;
    profile_columns(s, index, col, Lnu, Lnuc);
    if (model) *model = s->p_model[index];
    if (header) {
        header[0] = s->p_tsec[index] / k::julian_day; header[1] = s->Mcore; header[2] = s->Rcore; header[3] = s->p_Mdot[index];
        header[4] = Lnuc; header[5] = Lnu; header[6] = s->Mtotal; header[7] = s->Rtotal;
    }
    if (columns) std::copy(col.begin(), col.end(), columns);
    return 0;
}

// do_write_profile (NScool_profile.f:27-106): one file "<base><model>" per kept profile + the manifest
int dstar_NScool_write_profiles(int id, const char* directory) {
    Star* s = get(id);
    if (!s || !s->evolved || !directory) return -1;
    const std::string dir = directory, base = dir + "/" + s->c.base_profile_filename;
    FILE* mf = std::fopen((dir + "/" + s->c.profile_manifest_filename).c_str(), "w");
    if (!mf) { g_err = "write_profiles: cannot open manifest in " + dir; return -1; }
    std::fprintf(mf, "%15s%15s    %s\n", "model", "time [d]", ("base profile filename = " + s->c.base_profile_filename).c_str());
    static const char* hcols[9] = {"model", "time", "core_mass", "core_radius", "accretion_rate", "heating_lum", "cooling_lum",
                                   "total_mass", "total_radius"};
    static const char* pcols[20] = {"zone", "mass", "dm", "area", "gravity", "eLambda", "ePhi", "temperature", "luminosity",
                                    "pressure", "density", "zbar", "abar", "Xn", "Qimp", "Gamma", "cp", "eps_nu", "eps_nuc", "K"};
    for (size_t ip = 0; ip < s->p_model.size(); ++ip) {
        std::vector<double> col;
        double Lnu, Lnuc;
        profile_columns(s, (int)ip, col, Lnu, Lnuc);
        FILE* fp = std::fopen((base + std::to_string(s->p_model[ip])).c_str(), "w");
        if (!fp) { std::fclose(mf); g_err = "write_profiles: cannot open profile file"; return -1; }
        std::fprintf(fp, "dstar-b200/profile\n\n");
        for (int i = 1; i <= 9; ++i) std::fprintf(fp, "%15d", i);
        std::fprintf(fp, "\n");
        for (int i = 0; i < 9; ++i) std::fprintf(fp, "%15s", hcols[i]);
        std::fprintf(fp, "\n%15d%15.6E%15.6f%15.6f%15.6E%15.6E%15.6E%15.6f%15.6f\n\n\n", s->p_model[ip], s->p_tsec[ip] / k::julian_day,
                     s->Mcore, s->Rcore, s->p_Mdot[ip], Lnuc, Lnu, s->Mtotal, s->Rtotal);
        for (int i = 1; i <= 20; ++i) std::fprintf(fp, "%15d", i);
        std::fprintf(fp, "\n");
        for (int i = 0; i < 20; ++i) std::fprintf(fp, "%15s", pcols[i]);
        std::fprintf(fp, "\n");
        for (int iz = 0; iz < s->nz; ++iz) {
            const double* q = &col[(size_t)20 * iz];
            // (i15, 4es15.6, 2f15.6, 4es15.6, 5f15.6, 4es15.6)
            std::fprintf(fp, "%15d%15.6E%15.6E%15.6E%15.6E%15.6f%15.6f%15.6E%15.6E%15.6E%15.6E%15.6f%15.6f%15.6f%15.6f%15.6f%15.6E%15.6E%15.6E%15.6E\n",
                         iz + 1, q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10], q[11], q[12], q[13], q[14], q[15],
                         q[16], q[17], q[18], q[19]);
        }
        std::fclose(fp);
        std::fprintf(mf, "%15d%15.6E\n", s->p_model[ip], s->p_tsec[ip] / k::julian_day);
    }
    std::fclose(mf);
    return 0;
}

double dstar_NScool_last_kernel_ms(int which) { return (which >= 0 && which < 3) ? cur().ms[which] : -1.0; }

const char* dstar_NScool_last_error(void) { return g_err.c_str(); }

// ---- the hooks of examples/custom_Tc_Qimp/src/alt_micro.f, for sweep drivers
// alt_Qimp (:4-19): Q = extra_real_controls(2) wherever rho_bar > extra_real_controls(1)
void dstar_example_alt_Qimp(int id, int* ierr) {
    Star* s = get(id);
    *ierr = s ? 0 : -1;
    if (!s) return;
    const double rhoC = s->c.extra_real_controls[0], Q = s->c.extra_real_controls[1];
    auto& ib = s->z["ionic_bar"];
    const auto& rb = s->z["rho_bar"];
    for (int k2 = 0; k2 < s->nz; ++k2) if (rb[k2] > rhoC) ib[11 * k2 + 10] = Q;
}
// alt_sf (:21-44): Tc = 1e10 K everywhere, neutron 1S0 closes above the wavenumber of rho_C
void dstar_example_alt_sf(int id, double kp, double kn, double Tc[3]) {
    (void)kp;
    Tc[0] = Tc[1] = Tc[2] = (double)1.0e10f;
    Star* s = get(id);
    if (!s) return;
    const double rhoC = s->c.extra_real_controls[0];
    const double n = rhoC * k::avogadro;
    const double kC = dsm::pow(0.5 * k::threepisquare * n, k::onethird) / k::cm_to_fm;
    if (kn > kC) Tc[1] = 0.0;
}

}  // extern "C"
