// Seam B kernel, warp form: ONE WARP per star, Z contiguous zones per lane (nz <= 32 Z), many stars per SM.
//
// Replaces, per star, the epoch loop of NScool_evolve_model (NScool/public/NScool_lib.f:50-91) and
// do_integrate_crust (NScool/private/NScool_evolve.f:11-136): isolve(rodasp) with get_derivatives (:237-283),
// get_jacobian (:285-359), interpolate_temps (:452-459), get_coefficients (:461-524), get_nuclear_heating (:526-568),
// evaluate_luminosity (:433-450) and the bookkeeping of evaluate_timestep (:138-235).  Same formulas and the same
// RODASP driver as kernels_evolve.cuh (the CTA-per-star form, kept for nz > 320); what differs is the mapping:
//
//   * no CTA barrier anywhere in the time loop.  Lane l owns zones [l Z, l Z + Z); neighbour temperatures and
//     luminosities cross lanes by warp shuffle, norms are shuffle reductions.
//   * the tridiagonal Rosenbrock matrix (I/(gamma h) - J) is factorised once per step by a partition method: inside a
//     lane a downward and an upward elimination sweep (Z - 1 and Z - 2 steps, registers only) leave every row coupled
//     to just the last unknown of the previous lane and the last unknown of its own lane; the last rows of the 32
//     lanes form a 32 x 32 tridiagonal system that is solved by CYCLIC REDUCTION OVER WARP SHUFFLES (5 levels, all
//     multipliers kept in registers); the interior unknowns then follow without any further recurrence.  The six
//     stage systems of a step replay the stored multipliers: 2 Z dependent multiply-adds, 12 shuffles.
//     (The CTA form ran 8 levels of cyclic reduction through shared memory with a CTA barrier per level, 63 + 14
//     barriers per step, and held 2 stars per SM; this form holds 6.)
//   * per-zone constants, the five stage vectors and the Jacobian rows live in shared memory as [quantity][slot][lane]
//     (each lane reads its own column: conflict-free); ln T, the trial solution and the factorisation live in registers.
//
// A star's arithmetic depends on its own zone count only (the launch groups stars by Z class), so a star gives the
// same bits alone or inside any batch.
#pragma once
#include "kernels_evolve.cuh"

namespace dsw {
using namespace dsc;
constexpr unsigned FULL = 0xffffffffu;
enum { C_DM = 0, C_DMBAR, C_AREA, C_RHOBAR, C_EPHI, C_E2PHI, C_EPHIBAR, C_E2PHIBAR, C_ENUC, N_CONST };
constexpr int N_ATM = 12;   // per-star cache of the atmosphere-table interval zone 0 is in (lane 0)

template <int Z>
struct Lay {   // shared memory of one star, in doubles
    static constexpr int AK = 0;                          // [5][Z][32]  stage vectors ak1..ak5
    static constexpr int CONST = AK + 5 * Z * 32;         // [N_CONST][Z][32]
    static constexpr int JAC = CONST + N_CONST * Z * 32;  // [4][Z][32]  J(k,k-1), J(k,k), J(k,k+1), f(y)
    static constexpr int ATM = JAC + 4 * Z * 32;          // [N_ATM]
    static constexpr int TOTAL = ATM + N_ATM + 4;
};

DS_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
DS_D double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

template <int Z>
struct Lu {   // partition factorisation of one lane's rows + its row of the reduced system
    double m1[Z], m2[Z], a[Z], c[Z], binv[Z];
    double q, al[5], ga[5], Binv;
};

// rows (a, b, c) = (sub, diagonal, super) of this lane's Z zones; rows past the star's end are (0, 1, 0)
template <int Z>
DS_D void factor(int lane, double (&a)[Z], double (&b)[Z], double (&c)[Z], Lu<Z>& f) {
    // downward sweep: row i loses x_{i-1} and gains a coupling to x_left (the last unknown of the previous lane)
    f.m1[0] = 0.0;
#pragma unroll
    for (int i = 1; i < Z; ++i) {
        const double m = a[i] / b[i - 1];
        f.m1[i] = m;
        b[i] = b[i] - m * c[i - 1];
        a[i] = -m * a[i - 1];
    }
    // upward sweep over rows Z-3 .. 0: row i loses x_{i+1} and gains a coupling to x_last (this lane's last unknown)
#pragma unroll
    for (int i = 0; i < Z; ++i) f.m2[i] = 0.0;
#pragma unroll
    for (int i = Z - 3; i >= 0; --i) {
        const double m = c[i] / b[i + 1];
        f.m2[i] = m;
        a[i] = a[i] - m * a[i + 1];
        c[i] = -m * c[i + 1];
    }
    // reduced row of this lane: its last row, with the first unknown of the next lane eliminated through that
    // lane's first row  a0' u_l + b0' x_first' + c0' u_{l+1} = d0'
    const double a0n = __shfl_down_sync(FULL, a[0], 1), b0n = __shfl_down_sync(FULL, b[0], 1),
                 c0n = __shfl_down_sync(FULL, c[0], 1);
    f.q = (lane < 31) ? c[Z - 1] / b0n : 0.0;
    double A = a[Z - 1], B = b[Z - 1] - f.q * a0n, C = -f.q * c0n;
    if (lane == 31) { B = b[Z - 1]; C = 0.0; }
    // cyclic reduction over the 32 lanes
#pragma unroll
    for (int lev = 0; lev < 5; ++lev) {
        const int s = 1 << lev;
        const double Am = __shfl_up_sync(FULL, A, s), Bm = __shfl_up_sync(FULL, B, s), Cm = __shfl_up_sync(FULL, C, s);
        const double Ap = __shfl_down_sync(FULL, A, s), Bp = __shfl_down_sync(FULL, B, s), Cp = __shfl_down_sync(FULL, C, s);
        double al = 0.0, ga = 0.0, nA = 0.0, nC = 0.0;
        if (lane >= s) { al = -A / Bm; nA = al * Am; B = B + al * Cm; }
        if (lane + s < 32) { ga = -C / Bp; nC = ga * Cp; B = B + ga * Ap; }
        f.al[lev] = al; f.ga[lev] = ga;
        A = nA; C = nC;
    }
    f.Binv = 1.0 / B;
#pragma unroll
    for (int i = 0; i < Z; ++i) { f.a[i] = a[i]; f.c[i] = c[i]; f.binv[i] = 1.0 / b[i]; }
}
template <int Z>
DS_D void solve(const Lu<Z>& f, double (&d)[Z]) {   // d: right-hand side in, solution out
#pragma unroll
    for (int i = 1; i < Z; ++i) d[i] = d[i] - f.m1[i] * d[i - 1];
#pragma unroll
    for (int i = Z - 3; i >= 0; --i) d[i] = d[i] - f.m2[i] * d[i + 1];
    double D = d[Z - 1] - f.q * __shfl_down_sync(FULL, d[0], 1);
#pragma unroll
    for (int lev = 0; lev < 5; ++lev) {
        const int s = 1 << lev;
        const double Dm = __shfl_up_sync(FULL, D, s), Dp = __shfl_down_sync(FULL, D, s);
        D = D + f.al[lev] * Dm + f.ga[lev] * Dp;   // multipliers of lanes without that neighbour are zero
    }
    const double u = D * f.Binv;
    const double xl = __shfl_up_sync(FULL, u, 1);
#pragma unroll
    for (int i = 0; i < Z - 1; ++i) d[i] = (d[i] - f.a[i] * xl - f.c[i] * u) * f.binv[i];
    d[Z - 1] = u;
}

struct WCtx {
    int n, lane, z0;   // zones of the star, lane, first flattened zone
    int zc0, zf0;      // first slot of the star's cell / face tables (= z0 unless tables are shared)
    bool fixed_atm, insulating;
    int n_tab;
    double x0, inv_dx;
    const double* tab;   // lnT knots in shared memory
    int atm_nv;
    const double *atm_lgTb, *atm_lgTeff, *atm_lgflux;
    double atm_x0, atm_x1;
    const double *tab_lnK, *tab_lnCp, *tab_lnEnu;
};
struct Surface {   // meaningful in lane 0
    double Lsurf, dlgflux, Teff;
};

// get_derivatives (NScool_evolve.f:237-283) for the Z zones of this lane; JAC: also the rows of get_jacobian (:285-359)
// into shared memory, and sum_enu_dm = this lane's part of L_nu.
template <int Z, bool JAC>
DS_D void rhs(const WCtx& cx, double* sm, const double (&y)[Z], double (&f)[Z], Surface& sf, double& sum_enu_dm) {
    using L_ = Lay<Z>;
    const int lane = cx.lane, n = cx.n;
#define CZ(c_, i_) sm[L_::CONST + ((c_) * Z + (i_)) * 32 + lane]
    // constant of the zone before slot i / after slot i (the neighbouring lane's column at the block edges)
#define CZ_PREV(c_, i_) ((i_) > 0 ? sm[L_::CONST + ((c_) * Z + (i_) - 1) * 32 + lane] : sm[L_::CONST + ((c_) * Z + Z - 1) * 32 + lane - 1])
#define CZ_NEXT(c_, i_) ((i_) < Z - 1 ? sm[L_::CONST + ((c_) * Z + (i_) + 1) * 32 + lane] : sm[L_::CONST + ((c_) * Z) * 32 + lane + 1])
    double T[Z], L[Z], Kc[Z], dlnK[Z];
#pragma unroll
    for (int i = 0; i < Z; ++i) T[i] = (lane * Z + i < n) ? exp(y[i]) : 0.0;
    const double T_left = __shfl_up_sync(FULL, T[Z - 1], 1);
#pragma unroll
    for (int i = 0; i < Z; ++i) {
        const int k = lane * Z + i;
        L[i] = 0.0; Kc[i] = 0.0; dlnK[i] = 0.0;
        if (k < n) {
            // interpolate_temps :452-459
            const double Tm1 = (i > 0) ? T[i - 1] : T_left;
            const double T_bar = (k == 0) ? T[i] : 0.5 * (Tm1 * CZ(C_DM, i) + T[i] * CZ_PREV(C_DM, i)) / CZ(C_DMBAR, i);
            const double lnT_bar = log(T_bar);
            // get_coefficients :461-524
            const int kK = dsv::spline_locate(cx.tab, cx.n_tab, cx.x0, cx.inv_dx, lnT_bar);
            const double4 cK = reinterpret_cast<const double4*>(cx.tab_lnK + (size_t)4 * cx.n_tab * (cx.zf0 + k))[kK];
            double lnK;
            dsv::spline_poly(cK, lnT_bar - cx.tab[kK], lnK, dlnK[i]);
            Kc[i] = exp(lnK);
            // evaluate_luminosity :433-450
            if (k == 0) {
                double* at = sm + L_::ATM;
                const double lgTb = lnT_bar / ln10;
                const double lgTc = fmin(fmax(lgTb, cx.atm_x0), cx.atm_x1);
                const int aj = (int)at[0];
                const bool hit = aj >= 0 && lgTc >= at[1] && (lgTc < at[2] || (aj == cx.atm_nv - 2 && lgTc <= at[2]));
                if (!hit) {
                    const int j = ds_locate(cx.atm_lgTb, cx.atm_nv, lgTc);
                    at[0] = (double)j; at[1] = cx.atm_lgTb[j]; at[2] = cx.atm_lgTb[j + 1];
                    for (int q = 0; q < 4; ++q) { at[3 + q] = cx.atm_lgTeff[4 * j + q]; at[7 + q] = cx.atm_lgflux[4 * j + q]; }
                }
                const double dx = lgTc - at[1];
                const double* a4 = at + 3;
                const double* b4 = at + 7;
                const double lgTeff = a4[0] + dx * (a4[1] + dx * (a4[2] + dx * a4[3]));
                const double lgflux = b4[0] + dx * (b4[1] + dx * (b4[2] + dx * b4[3]));
                sf.dlgflux = b4[1] + 2.0 * dx * (b4[2] + 1.5 * dx * b4[3]);
                L[i] = CZ(C_AREA, i) * dsd::dexp10(lgflux);
                sf.Lsurf = L[i];
                sf.Teff = dsd::dexp10(lgTeff);
            } else {
                L[i] = -(CZ(C_AREA, i) * CZ(C_AREA, i)) * CZ(C_RHOBAR, i) * Kc[i] * (CZ_PREV(C_EPHI, i) * Tm1 - CZ(C_EPHI, i) * T[i]) /
                       CZ(C_EPHIBAR, i) / CZ(C_DMBAR, i);
            }
        }
    }
    const double L_right = __shfl_down_sync(FULL, L[0], 1);
    // Jacobian pieces that a row needs from the zone after it
    double dLm[Z], dLpF[Z];
    if (JAC) {
#pragma unroll
        for (int i = 0; i < Z; ++i) {
            const int k = lane * Z + i;
            dLm[i] = 0.0; dLpF[i] = 0.0;
            if (k < n && k > 0) {
                const double Tm1 = (i > 0) ? T[i - 1] : T_left;
                const double dm = CZ(C_DM, i), dm_m1 = CZ_PREV(C_DM, i);
                const double ArKdm = (CZ(C_AREA, i) * CZ(C_AREA, i)) * CZ(C_RHOBAR, i) * Kc[i] / CZ(C_EPHIBAR, i) / CZ(C_DMBAR, i);
                const double den = dm * Tm1 + dm_m1 * T[i];
                const double wplus = dm * Tm1 / den;
                const double wminus = dm_m1 * T[i] / den;
                dLpF[i] = L[i] * dlnK[i] * wplus - ArKdm * CZ_PREV(C_EPHI, i) * Tm1;
                dLm[i] = L[i] * dlnK[i] * wminus + ArKdm * CZ(C_EPHI, i) * T[i];
            }
        }
    }
    const double dLm_right = JAC ? __shfl_down_sync(FULL, dLm[0], 1) : 0.0;
    const double dLpF_right = JAC ? __shfl_down_sync(FULL, dLpF[0], 1) : 0.0;
    sum_enu_dm = 0.0;
#pragma unroll
    for (int i = 0; i < Z; ++i) {
        const int k = lane * Z + i;
        f[i] = 0.0;
        double lo = 0.0, di = 0.0, up = 0.0;
        if (k < n) {
            const int kY = dsv::spline_locate(cx.tab, cx.n_tab, cx.x0, cx.inv_dx, y[i]);
            const double4 cCp = reinterpret_cast<const double4*>(cx.tab_lnCp + (size_t)4 * cx.n_tab * (cx.zc0 + k))[kY];
            const double4 cEn = reinterpret_cast<const double4*>(cx.tab_lnEnu + (size_t)4 * cx.n_tab * (cx.zc0 + k))[kY];
            double lnCp, dlnCp, lnenu, dlnenu;
            dsv::spline_poly(cCp, y[i] - cx.tab[kY], lnCp, dlnCp);
            dsv::spline_poly(cEn, y[i] - cx.tab[kY], lnenu, dlnenu);
            const double Cp = exp(lnCp), enu = exp(lnenu);
            const double dm = CZ(C_DM, i), e2Phi = CZ(C_E2PHI, i), ePhi = CZ(C_EPHI, i), e2Phi_bar = CZ(C_E2PHIBAR, i);
            const double Lp1 = (i < Z - 1) ? L[i + 1] : L_right;
            const double e2Phi_bar_p1 = (k < n - 1) ? CZ_NEXT(C_E2PHIBAR, i) : 0.0;
            double fk = 0.0;
            if (k < n - 1)
                fk = ((Lp1 * e2Phi_bar_p1 - L[i] * e2Phi_bar) / dm / e2Phi + CZ(C_ENUC, i) - enu) * ePhi / Cp / T[i];
            else if (cx.insulating)
                fk = ((-L[i] * e2Phi_bar) / dm / e2Phi + CZ(C_ENUC, i) - enu) * ePhi / Cp / T[i];
            if (cx.fixed_atm && k == 0) fk = 0.0;
            f[i] = fk;
            if (JAC) {
                sum_enu_dm += enu * dm;
                const double CTinv = ePhi / (Cp * T[i]);
                const double CTdminv = CTinv / e2Phi / dm;
                const double dLm_p1 = (i < Z - 1) ? dLm[i + 1] : dLm_right;
                const double dLpF_p1 = (i < Z - 1) ? dLpF[i + 1] : dLpF_right;
                di = -fk * (1.0 + dlnCp) - CTinv * enu * dlnenu;
                if (k > 0 && k < n - 1) di = di + CTdminv * (e2Phi_bar_p1 * dLpF_p1 - e2Phi_bar * dLm[i]);
                if (k == 0) di = di + CTdminv * (e2Phi_bar_p1 * dLpF_p1 - e2Phi_bar * L[i] * sf.dlgflux);
                if (k == n - 1) {
                    if (cx.insulating) di = di + CTdminv * (-e2Phi_bar * dLm[i]);
                    else di = 0.0;
                }
                if (k < n - 1) up = CTdminv * e2Phi_bar_p1 * dLm_p1;
                if (k > 0 && k < n - 1) lo = -CTdminv * e2Phi_bar * dLpF[i];
                if (cx.fixed_atm) {
                    if (k == 0) { up = 0.0; di = 0.0; }
                    if (k == 1) lo = 0.0;
                }
            }
        }
        if (JAC) {
            sm[L_::JAC + (0 * Z + i) * 32 + lane] = lo;
            sm[L_::JAC + (1 * Z + i) * 32 + lane] = di;
            sm[L_::JAC + (2 * Z + i) * 32 + lane] = up;
            sm[L_::JAC + (3 * Z + i) * 32 + lane] = f[i];
        }
    }
#undef CZ
#undef CZ_PREV
#undef CZ_NEXT
}
}  // namespace dsw

// S stars (warps) per CTA; star_list: the stars of this Z class
template <int Z, int S>
__global__ void __launch_bounds__(32 * S) ds_evolve_warp_kernel(DsEvolveArgs a, const int* __restrict__ star_list, int n_list) {
    using namespace dsw;
    using namespace dsv;   // RODASP coefficients
    using L_ = Lay<Z>;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* tab = smem;                               // (n_tab) lnT knots, shared by the CTA's stars
    double* sm = smem + DS_NTAB + (size_t)w * L_::TOTAL;
    for (int i = threadIdx.x; i < a.n_tab; i += blockDim.x) tab[i] = a.tab_lnT[i];
    __syncthreads();   // the only CTA barrier
    const int li = blockIdx.x * S + w;
    if (li >= n_list) return;
    const int star = star_list[li];
    WCtx cx;
    cx.lane = lane;
    cx.z0 = a.zone_offset[star];
    cx.n = a.zone_offset[star + 1] - cx.z0;
    cx.zc0 = a.cell_tab_zone0 ? a.cell_tab_zone0[star] : cx.z0;
    cx.zf0 = a.face_tab_zone0 ? a.face_tab_zone0[star] : cx.z0;
    const int n = cx.n, z0 = cx.z0;
    cx.n_tab = a.n_tab; cx.tab = tab;
    cx.x0 = tab[0];
    cx.inv_dx = (double)(a.n_tab - 1) / (tab[a.n_tab - 1] - tab[0]);
    cx.insulating = a.make_inner_boundary_insulating && !a.fix_core_temperature;
    const int ai = a.atm_index[star];
    cx.atm_nv = a.atm_nv;
    cx.atm_lgTb = a.atm_lgTb + (size_t)ai * a.atm_nv;
    cx.atm_lgTeff = a.atm_lgTeff + (size_t)ai * 4 * a.atm_nv;
    cx.atm_lgflux = a.atm_lgflux + (size_t)ai * 4 * a.atm_nv;
    cx.atm_x0 = cx.atm_lgTb[0]; cx.atm_x1 = cx.atm_lgTb[a.atm_nv - 1];
    cx.tab_lnK = a.tab_lnK; cx.tab_lnCp = a.tab_lnCp; cx.tab_lnEnu = a.tab_lnEnu;
    if (lane == 0) sm[L_::ATM] = -1.0;
#define CZ(c_, i_) sm[L_::CONST + ((c_) * Z + (i_)) * 32 + lane]
#define AK(j_, i_) sm[L_::AK + ((j_) * Z + (i_)) * 32 + lane]
#define JC(c_, i_) sm[L_::JAC + ((c_) * Z + (i_)) * 32 + lane]
    double y[Z], Pz[Z];
#pragma unroll
    for (int i = 0; i < Z; ++i) {
        const int k = lane * Z + i, g = z0 + k;
        const bool on = k < n;
        CZ(C_DM, i) = on ? a.dm[g] : 1.0; CZ(C_DMBAR, i) = on ? a.dm_bar[g] : 1.0; CZ(C_AREA, i) = on ? a.area[g] : 0.0;
        CZ(C_RHOBAR, i) = on ? a.rho_bar[g] : 0.0; CZ(C_EPHI, i) = on ? a.ePhi[g] : 1.0; CZ(C_E2PHI, i) = on ? a.e2Phi[g] : 1.0;
        CZ(C_EPHIBAR, i) = on ? a.ePhi_bar[g] : 1.0; CZ(C_E2PHIBAR, i) = on ? a.e2Phi_bar[g] : 1.0; CZ(C_ENUC, i) = 0.0;
        Pz[i] = on ? a.P[g] : 0.0;
        y[i] = on ? a.lnT[g] : 0.0;
    }
    __syncwarp();
    const double zmin = a.min_lg_T * ln10, zmax = a.max_lg_T * ln10;
    const double uround = 1.0e-16, fac1 = 5.0, fac2 = 1.0 / 6.0, safe = 0.9;
    double dt_state = a.dt[star], tsec = a.tsec[star];
    int model = a.model[star];
    int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    int start_num = a.starting_number_for_profile;
    int ntrace = 0, nprof = 0;
    Surface sf{0.0, 0.0, 0.0};
    double Teff = 0.0, Lsurf = 0.0;

    for (int ep = 0; ep < a.n_epoch; ++ep) {
        const double t_start = a.epoch_boundaries[(size_t)ep * a.n_star + star];
        const double duration = a.epoch_boundaries[(size_t)(ep + 1) * a.n_star + star] - t_start;
        const double Mdot = a.epoch_Mdots[(size_t)ep * a.n_star + star];
        if (ep > 0) start_num = a.suppress_first_step_output ? model : model + 1;  // NScool_lib.f:72-78
        cx.fixed_atm = a.fix_atmosphere_temperature_when_accreting && Mdot > 0.0;
        // get_nuclear_heating (NScool_evolve.f:526-568): later windows overwrite earlier ones
        {
            double enuc[Z];
#pragma unroll
            for (int i = 0; i < Z; ++i) enuc[i] = 0.0;
            const double* hw = a.heat + (size_t)9 * star;
            const int nwin = a.turn_on_extra_heating ? 3 : 2;
            for (int win = 0; win < nwin; ++win) {
                const double Ptop = dsd::dexp10(hw[3 * win]), Pbot = dsd::dexp10(hw[3 * win + 1]), Q = hw[3 * win + 2];
                double part = 0.0;
#pragma unroll
                for (int i = 0; i < Z; ++i)
                    if (lane * Z + i < n && Pz[i] >= Ptop && Pz[i] <= Pbot) part += CZ(C_DM, i);
                const double M_norm = warp_sum(part);
#pragma unroll
                for (int i = 0; i < Z; ++i)
                    if (lane * Z + i < n && Pz[i] >= Ptop && Pz[i] <= Pbot) enuc[i] = Mdot * Q * mev_to_ergs * avogadro / M_norm;
            }
#pragma unroll
            for (int i = 0; i < Z; ++i) {
                if (a.enuc_override && lane * Z + i < n) enuc[i] = a.enuc_override[z0 + lane * Z + i] * Mdot;
                CZ(C_ENUC, i) = enuc[i];
            }
        }
        if (cx.fixed_atm && lane == 0) y[0] = log(a.T_atm[star]);  // NScool_evolve.f:69-71
        const double rtol = a.integration_tolerance;
        double ymin = 1.0e300;
#pragma unroll
        for (int i = 0; i < Z; ++i)
            if (lane * Z + i < n) ymin = fmin(ymin, y[i]);
        const double atol = a.integration_tolerance * warp_min(ymin);

        // ---- RODAS driver (Hairer & Wanner), RODASP coefficients
        double x = 0.0, xend = duration;
        const double posneg = copysign(1.0, xend - x);
        double hmax = (a.maximum_timestep == 0.0) ? (xend - x) : a.maximum_timestep;
        hmax = fmin(fabs(hmax), fabs(xend - x));
        double h = dt_state;
        if (fabs(h) <= 10.0 * uround) h = 1.0e-6;
        h = fmin(fabs(h), hmax);
        h = copysign(h, posneg);
        bool reject = false, last = false;
        int idid = 1, naccpt = 0, nstep = 0;   // nstep: isolve's NSTEP starts at 0 for every epoch (NScool_evolve.f:111)
        double hacc = 0.0, erracc = 0.0, hopt = h, xold = x;
        bool first = true;
        for (;;) {
            // f(x, y) and the Jacobian rows at the current accepted point (also evaluate_timestep's refresh)
            double fy[Z], lnu_part;
            rhs<Z, true>(cx, sm, y, fy, sf, lnu_part);
            Teff = sf.Teff; Lsurf = sf.Lsurf;
            if (!(first && a.suppress_first_step_output)) {
                model = naccpt + 1 + start_num;
                tsec = x + t_start;
                dt_state = x - xold;
                if (a.trace_cap > 0) {
                    double pn = 0.0;
#pragma unroll
                    for (int i = 0; i < Z; ++i) pn += CZ(C_ENUC, i) * ((lane * Z + i < n) ? CZ(C_DM, i) : 0.0);
                    const double Lnu = warp_sum(lnu_part), Lnuc = warp_sum(pn);
                    if (lane == 0 && ntrace < a.trace_cap) {
                        const size_t r = (size_t)ntrace * a.n_star + star;
                        a.trace_model[r] = model; a.trace_tsec[r] = tsec; a.trace_dt[r] = dt_state;
                        a.trace_Teff[r] = Teff; a.trace_Lsurf[r] = Lsurf; a.trace_Lnu[r] = Lnu;
                        a.trace_Lnuc[r] = Lnuc; a.trace_Mdot[r] = Mdot;
                    }
                    {   // the terminal line's extrema (NScool_terminal.f:53-56)
                        double ymax = -1.0e300, ymn = 1.0e300;
#pragma unroll
                        for (int i = 0; i < Z; ++i)
                            if (lane * Z + i < n) { ymax = fmax(ymax, y[i]); ymn = fmin(ymn, y[i]); }
                        ymax = -warp_min(-ymax); ymn = warp_min(ymn);
                        double kx = 1.0e9, kn = 1.0e9;
#pragma unroll
                        for (int i = Z - 1; i >= 0; --i)
                            if (lane * Z + i < n) {
                                if (y[i] == ymax) kx = (double)(lane * Z + i);
                                if (y[i] == ymn) kn = (double)(lane * Z + i);
                            }
                        const int kmax = (int)warp_min(kx), kmin = (int)warp_min(kn);
                        if (ntrace < a.trace_cap) {
                            const size_t cs = (size_t)a.trace_cap * a.n_star, r = (size_t)ntrace * a.n_star + star;
#pragma unroll
                            for (int i = 0; i < Z; ++i) {
                                if (lane * Z + i == kmax) { a.trace_ext[r] = ymax; a.trace_ext[cs + r] = Pz[i]; }
                                if (lane * Z + i == kmin) { a.trace_ext[2 * cs + r] = ymn; a.trace_ext[3 * cs + r] = Pz[i]; }
                            }
                        }
                    }
                    ++ntrace;
                }
                if (a.prof_cap > 0 && model % a.prof_interval == 0) {
                    if (nprof < a.prof_cap) {
#pragma unroll
                        for (int i = 0; i < Z; ++i)
                            if (lane * Z + i < n) a.prof_lnT[(size_t)nprof * a.total_zones + z0 + lane * Z + i] = y[i];
                        if (lane == 0) {
                            const size_t r = (size_t)nprof * a.n_star + star;
                            a.prof_model[r] = model; a.prof_tsec[r] = tsec; a.prof_Mdot[r] = Mdot;
                        }
                    }
                    ++nprof;
                }
            }
            first = false;
            if (nstep > a.maximum_number_of_models) { idid = -2; break; }
            if (0.1 * fabs(h) <= fabs(x) * uround) { idid = -3; break; }
            if (last) { h = hopt; idid = 1; break; }
            hopt = h;
            if ((x + h * 1.0001 - xend) * posneg >= 0.0) { h = xend - x; last = true; }
            ++cnt[0]; ++cnt[1];
            bool accepted = false;
            while (!accepted) {
                const double fac = 1.0 / (h * GAMMA);
                Lu<Z> lu;
                {
                    double ra[Z], rb[Z], rc[Z];
#pragma unroll
                    for (int i = 0; i < Z; ++i) {
                        const bool on = lane * Z + i < n;
                        ra[i] = on ? -JC(0, i) : 0.0; rb[i] = on ? fac - JC(1, i) : 1.0; rc[i] = on ? -JC(2, i) : 0.0;
                    }
                    factor<Z>(lane, ra, rb, rc, lu);
                }
                ++cnt[5];
                const double hc21 = C21 / h, hc31 = C31 / h, hc32 = C32 / h, hc41 = C41 / h, hc42 = C42 / h,
                             hc43 = C43 / h, hc51 = C51 / h, hc52 = C52 / h, hc53 = C53 / h, hc54 = C54 / h,
                             hc61 = C61 / h, hc62 = C62 / h, hc63 = C63 / h, hc64 = C64 / h, hc65 = C65 / h;
                double ak[Z], ynew[Z], fs[Z], dummy;
                // stage 1
#pragma unroll
                for (int i = 0; i < Z; ++i) ak[i] = JC(3, i);
                solve<Z>(lu, ak);
#pragma unroll
                for (int i = 0; i < Z; ++i) { AK(0, i) = ak[i]; ynew[i] = y[i] + A21 * ak[i]; }
                rhs<Z, false>(cx, sm, ynew, fs, sf, dummy);
                // stage 2
#pragma unroll
                for (int i = 0; i < Z; ++i) ak[i] = fs[i] + hc21 * AK(0, i);
                solve<Z>(lu, ak);
#pragma unroll
                for (int i = 0; i < Z; ++i) { AK(1, i) = ak[i]; ynew[i] = y[i] + A31 * AK(0, i) + A32 * ak[i]; }
                rhs<Z, false>(cx, sm, ynew, fs, sf, dummy);
                // stage 3
#pragma unroll
                for (int i = 0; i < Z; ++i) ak[i] = fs[i] + (hc31 * AK(0, i) + hc32 * AK(1, i));
                solve<Z>(lu, ak);
#pragma unroll
                for (int i = 0; i < Z; ++i) { AK(2, i) = ak[i]; ynew[i] = y[i] + A41 * AK(0, i) + A42 * AK(1, i) + A43 * ak[i]; }
                rhs<Z, false>(cx, sm, ynew, fs, sf, dummy);
                // stage 4
#pragma unroll
                for (int i = 0; i < Z; ++i) ak[i] = fs[i] + (hc41 * AK(0, i) + hc42 * AK(1, i) + hc43 * AK(2, i));
                solve<Z>(lu, ak);
#pragma unroll
                for (int i = 0; i < Z; ++i) {
                    AK(3, i) = ak[i];
                    ynew[i] = y[i] + A51 * AK(0, i) + A52 * AK(1, i) + A53 * AK(2, i) + A54 * ak[i];
                }
                rhs<Z, false>(cx, sm, ynew, fs, sf, dummy);
                // stage 5
#pragma unroll
                for (int i = 0; i < Z; ++i) ak[i] = fs[i] + (hc52 * AK(1, i) + hc54 * AK(3, i) + hc51 * AK(0, i) + hc53 * AK(2, i));
                solve<Z>(lu, ak);
#pragma unroll
                for (int i = 0; i < Z; ++i) { AK(4, i) = ak[i]; ynew[i] = ynew[i] + ak[i]; }
                rhs<Z, false>(cx, sm, ynew, fs, sf, dummy);
                // stage 6
#pragma unroll
                for (int i = 0; i < Z; ++i)
                    ak[i] = fs[i] + (hc61 * AK(0, i) + hc62 * AK(1, i) + hc65 * AK(4, i) + hc64 * AK(3, i) + hc63 * AK(2, i));
                solve<Z>(lu, ak);
                cnt[6] += 6; cnt[0] += 5; ++cnt[2]; ++nstep;
                // error estimate, bounds
                double e2 = 0.0;
                int bad = 0;
#pragma unroll
                for (int i = 0; i < Z; ++i) {
                    ynew[i] = ynew[i] + ak[i];
                    if (lane * Z + i < n) {
                        const double sk = atol + rtol * fmax(fabs(y[i]), fabs(ynew[i]));
                        e2 += (ak[i] / sk) * (ak[i] / sk);
                        bad |= !(ynew[i] >= zmin && ynew[i] <= zmax);
                    }
                }
                const double err = sqrt(warp_sum(e2) / n);
                bad = __any_sync(FULL, bad || !(err == err));
                if (bad) {  // out of [zmin,zmax] or not a number: redo the step with h/2 until the step-size floor ends it
                    h = h * 0.5; reject = true; last = false;
                    if (naccpt >= 1) ++cnt[4];
                    if (nstep > a.maximum_number_of_models) { idid = -2; break; }
                    if (0.1 * fabs(h) <= fabs(x) * uround) { idid = -3; break; }
                    continue;
                }
                double facs = fmax(fac2, fmin(fac1, dsm::pow(err, 0.25) / safe));
                double hnew = h / facs;
                if (err <= 1.0) {
                    ++naccpt; ++cnt[3];
                    if (naccpt > 1) {
                        double facgus = (hacc / h) * dsm::pow(err * err / erracc, 0.25) / safe;
                        facgus = fmax(fac2, fmin(fac1, facgus));
                        facs = fmax(facs, facgus);
                        hnew = h / facs;
                    }
                    hacc = h;
                    erracc = fmax(1.0e-2, err);
#pragma unroll
                    for (int i = 0; i < Z; ++i) y[i] = ynew[i];
                    xold = x;
                    x = x + h;
                    if (fabs(hnew) > hmax) hnew = posneg * hmax;
                    if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
                    reject = false;
                    h = hnew;
                    accepted = true;
                } else {
                    reject = true; last = false;
                    h = hnew;
                    if (naccpt >= 1) ++cnt[4];
                    if (nstep > a.maximum_number_of_models) { idid = -2; break; }
                    if (0.1 * fabs(h) <= fabs(x) * uround) { idid = -3; break; }
                }
            }
            if (idid < 0) break;
        }
        if (lane == 0) {
            const size_t r = (size_t)ep * a.n_star + star;
            a.idid[r] = idid;
            if (idid >= 0) {   // the reference leaves the epoch loop before setting the monitors (NScool_lib.f:80-89)
                a.t_monitor[r] = tsec / julian_day;                   // NScool_lib.f:82
                a.Teff_monitor[r] = Teff * CZ(C_EPHI, 0);             // :83 (ePhi(1): zone 0 is the surface)
                a.Qb_monitor[r] = (Mdot > 0.0) ? Lsurf / Mdot * amu * ergs_to_mev : 0.0;  // :85-89
            }
        }
        if (idid < 0) {  // integration error: exit the loop over epochs (NScool_lib.f:81)
            for (int e2 = ep + 1; e2 < a.n_epoch; ++e2)
                if (lane == 0) a.idid[(size_t)e2 * a.n_star + star] = 0;
            break;
        }
    }
#pragma unroll
    for (int i = 0; i < Z; ++i)
        if (lane * Z + i < n) a.lnT[z0 + lane * Z + i] = y[i];
    if (lane == 0) {
        a.dt[star] = dt_state; a.tsec[star] = tsec; a.model[star] = model;
        a.Teff[star] = Teff; a.Lsurf[star] = Lsurf;
        for (int i = 0; i < 7; ++i) a.counters[(size_t)7 * star + i] = cnt[i];
        if (a.trace_count) a.trace_count[star] = ntrace;
        if (a.prof_count) a.prof_count[star] = nprof;
    }
#undef CZ
#undef AK
#undef JC
}
