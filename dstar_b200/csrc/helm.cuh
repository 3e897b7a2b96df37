// Device-side electron-positron Helmholtz table lookup (biquintic/bicubic Hermite).
//
// Replaces, for the NScool hot path, get_helm_eos_results (dStar_eos/private/electron.f:10-51)
// -> helmeos2 (helm_core.f:12-158) -> helmeos2aux (:161-503) + helm_electron_positron.dek.
// Only the seven quantities NScool consumes are produced (electron.f:43-50):
//   eele, pele, sele, deept (cv), dpepd, dpept, etaele.
//
// HBM layout: one 17-double record per table node (i,j), i fastest:
//   [f ft ftt fd fdd fdt fddt fdtt fddtt | dpdf dpdft dpdfd dpdfdt | ef eft efd efdt]
// so the four corners of a cell are two contiguous 272-byte runs instead of 68 scattered loads.
//
// The interpolant differences large, nearly equal free-energy values (e.g. f(j)-f(j+1) inside
// df/dT), so its result carries roundoff noise far above 1e-10 in the degenerate crust.  To be
// identical to the reference arithmetic the sums below are evaluated in the reference's term order
// with individually rounded multiplies and adds (no FMA contraction): type Sd.
#pragma once
#include "dconst.cuh"

struct DsHelmDev {
    int imax, jmax;
    double logtlo, logtstpi, logdlo, logdstpi, templo, temphi, logthi, denlo, denhi;
    const double* node;  // (17, imax, jmax)
    const double* t;     // (jmax)
    const double* d;     // (imax)
};
constexpr int DS_HELM_REC = 17;

namespace dsh {
// strictly rounded double: every operation is one IEEE op, never contracted
struct Sd {
    double v;
};
DS_D Sd operator*(Sd a, Sd b) { return Sd{__dmul_rn(a.v, b.v)}; }
DS_D Sd operator+(Sd a, Sd b) { return Sd{__dadd_rn(a.v, b.v)}; }
DS_D Sd operator-(Sd a, Sd b) { return Sd{__dsub_rn(a.v, b.v)}; }
DS_D Sd operator-(Sd a) { return Sd{-a.v}; }
DS_D Sd S(double x) { return Sd{x}; }

// quintic Hermite basis (helm_core.f:222-241)
DS_D Sd psi0(Sd z) { return z * z * z * (z * (S(-6.0) * z + S(15.0)) - S(10.0)) + S(1.0); }
DS_D Sd dpsi0(Sd z) { return z * z * (z * (S(-30.0) * z + S(60.0)) - S(30.0)); }
DS_D Sd ddpsi0(Sd z) { return z * (z * (S(-120.0) * z + S(180.0)) - S(60.0)); }
DS_D Sd psi1(Sd z) { return z * (z * z * (z * (S(-3.0) * z + S(8.0)) - S(6.0)) + S(1.0)); }
DS_D Sd dpsi1(Sd z) { return z * z * (z * (S(-15.0) * z + S(32.0)) - S(18.0)) + S(1.0); }
DS_D Sd ddpsi1(Sd z) { return z * (z * (S(-60.0) * z + S(96.0)) - S(36.0)); }
DS_D Sd psi2(Sd z) { return S(0.5) * z * z * (z * (z * (-z + S(3.0)) - S(3.0)) + S(1.0)); }
DS_D Sd dpsi2(Sd z) { return S(0.5) * z * (z * (z * (S(-5.0) * z + S(12.0)) - S(9.0)) + S(2.0)); }
DS_D Sd ddpsi2(Sd z) { return S(0.5) * (z * (z * (S(-20.0) * z + S(36.0)) - S(18.0)) + S(2.0)); }
// cubic Hermite basis (helm_core.f:268-278)
DS_D Sd xpsi0(Sd z) { return z * z * (S(2.0) * z - S(3.0)) + S(1.0); }
DS_D Sd xpsi1(Sd z) { return z * (z * (z - S(2.0)) + S(1.0)); }

struct W6 {
    Sd w0, w1, w2, w0m, w1m, w2m;
};
struct W4 {
    Sd w0, w1, w0m, w1m;
};
// corner order in the reference's fi(): (iat,jat), (iat+1,jat), (iat,jat+1), (iat+1,jat+1)
struct Corners {
    const double *c00, *c10, *c01, *c11;
};
#define FQ(k, A) S(__ldg(q.A + (k)))
// Biquintic sums in the reference's term order (helm_core.f:244-262): for the record fields k = 0..8
// (f, ft, ftt, fd, fdd, fdt, fddt, fdtt, fddtt) the density weight w0/w1/w2 and the temperature weight w0/w1/w2 that
// multiply it, the four corners of each field in the order (i,j), (i+1,j), (i,j+1), (i+1,j+1):
//   sum = f00 d.w t.w + f10 d.wm t.w + f01 d.w t.wm + f11 d.wm t.wm + (next field) ...   accumulated left to right.
// h5n forms NT such sums over the SAME density weights d and NT temperature-weight sets, term by term: the products
// f * d.w are formed once and used NT times.  (Written as five independent sums, the compiler merged them the other
// way round: all 72 products f * sd.w, f * dsd.w of the first sums were kept for the later ones -- a 580-byte spill
// frame per thread.)
template <int I> DS_D Sd wP(const W6& w) { return I == 0 ? w.w0 : I == 1 ? w.w1 : w.w2; }
template <int I> DS_D Sd wM(const W6& w) { return I == 0 ? w.w0m : I == 1 ? w.w1m : w.w2m; }
template <int NT, int K, int DI, int TI, bool FIRST>
DS_D void h5n_field(const Corners& q, const W6& d, const W6* const (&t)[NT], Sd (&r)[NT]) {
    const Sd p00 = FQ(K, c00) * wP<DI>(d), p10 = FQ(K, c10) * wM<DI>(d);
    const Sd p01 = FQ(K, c01) * wP<DI>(d), p11 = FQ(K, c11) * wM<DI>(d);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const Sd tp = wP<TI>(*t[n]), tm = wM<TI>(*t[n]);
        const Sd a = FIRST ? p00 * tp + p10 * tp : r[n] + p00 * tp + p10 * tp;
        r[n] = a + p01 * tm + p11 * tm;
    }
}
template <int NT>
DS_D void h5n(const Corners& q, const W6& d, const W6* const (&t)[NT], Sd (&r)[NT]) {
    h5n_field<NT, 0, 0, 0, true>(q, d, t, r);
    h5n_field<NT, 1, 0, 1, false>(q, d, t, r);
    h5n_field<NT, 2, 0, 2, false>(q, d, t, r);
    h5n_field<NT, 3, 1, 0, false>(q, d, t, r);
    h5n_field<NT, 4, 2, 0, false>(q, d, t, r);
    h5n_field<NT, 5, 1, 1, false>(q, d, t, r);
    h5n_field<NT, 6, 2, 1, false>(q, d, t, r);
    h5n_field<NT, 7, 1, 2, false>(q, d, t, r);
    h5n_field<NT, 8, 2, 2, false>(q, d, t, r);
}
// bicubic sum (helm_core.f:282-290); base = first field of the 4-field group
DS_D Sd h3(const Corners& q, int base, const W4& t, const W4& d) {
    Sd r = FQ(base, c00) * d.w0 * t.w0 + FQ(base, c10) * d.w0m * t.w0;
    r = r + FQ(base, c01) * d.w0 * t.w0m + FQ(base, c11) * d.w0m * t.w0m;
    r = r + FQ(base + 1, c00) * d.w0 * t.w1 + FQ(base + 1, c10) * d.w0m * t.w1;
    r = r + FQ(base + 1, c01) * d.w0 * t.w1m + FQ(base + 1, c11) * d.w0m * t.w1m;
    r = r + FQ(base + 2, c00) * d.w1 * t.w0 + FQ(base + 2, c10) * d.w1m * t.w0;
    r = r + FQ(base + 2, c01) * d.w1 * t.w0m + FQ(base + 2, c11) * d.w1m * t.w0m;
    r = r + FQ(base + 3, c00) * d.w1 * t.w1 + FQ(base + 3, c10) * d.w1m * t.w1;
    r = r + FQ(base + 3, c01) * d.w1 * t.w1m + FQ(base + 3, c11) * d.w1m * t.w1m;
#undef FQ
    return r;
}
}  // namespace dsh

struct DsHelmOut {
    double eele, pele, sele, deept, dpepd, dpept, etaele;
};

// Density-direction half of the interpolation: rho*Ye, its table cell and the 16 basis-function weights depend on
// the zone only, not on T -- computed once per zone (ds_zone_aux_kernel) and shared by its 128 temperature lanes.
struct DsHelmZone {
    double ye, din;
    int i0, err;   // err = -3: rho*Ye below the table
    dsh::W6 sd, dsd;
    dsh::W4 cd;
};
DS_D DsHelmZone ds_helm_zone(const DsHelmDev& h, double rho, double abar, double zbar) {
    using namespace dsh;
    DsHelmZone z{};
    const double ytot1 = 1.0 / abar;
    z.ye = __dmul_rn(ytot1, zbar);
    double din = __dmul_rn(z.ye, rho);
    z.err = (din < h.denlo) ? -3 : 0;
    if (din > h.denhi) din = h.denhi;
    z.din = din;
    if (z.err) return z;
    int iat = (int)((dsm::log10(din) - h.logdlo) * h.logdstpi) + 1;
    iat = max(1, min(iat, h.imax - 1));
    z.i0 = iat - 1;
    const double di = __ldg(h.d + z.i0), di1 = __ldg(h.d + z.i0 + 1);
    // helm_alloc.f:301-326 (same operations as the *_sav arrays)
    const Sd dd = S(di1) - S(di);
    const Sd dd2 = dd * dd;
    const Sd ddi = S(__ddiv_rn(1.0, dd.v));
    Sd xd = (S(din) - S(di)) * ddi;
    if (xd.v < 0.0) xd.v = 0.0;
    const Sd mxd = S(1.0) - xd;
    z.sd = W6{psi0(xd), psi1(xd) * dd, psi2(xd) * dd2, psi0(mxd), -psi1(mxd) * dd, psi2(mxd) * dd2};
    z.dsd = W6{dpsi0(xd) * ddi, dpsi1(xd), dpsi2(xd) * dd, -dpsi0(mxd) * ddi, dpsi1(mxd), -dpsi2(mxd) * dd};
    z.cd = W4{xpsi0(xd), xpsi1(xd) * dd, xpsi0(mxd), -xpsi1(mxd) * dd};
    return z;
}

// returns 0 ok, -1 non-positive entropy, -2 logT < 5 (blend region, never reached by NScool),
// -3 rho*Ye below the table
DS_D int ds_helm_electron(const DsHelmDev& h, double T, double logT, const DsHelmZone& hz, DsHelmOut& o) {
    using namespace dsh;
    o = DsHelmOut{0, 0, 0, 0, 0, 0, 0};
    if (!(logT >= 5.0)) return -2;
    double temp = T, logtemp = logT;
    const double ye = hz.ye, din = hz.din;
    if (temp < h.templo) { temp = h.templo; logtemp = dsm::log10(temp); }
    if (hz.err) return hz.err;
    if (temp > h.temphi) { temp = h.temphi; logtemp = h.logthi; }

    int jat = (int)((logtemp - h.logtlo) * h.logtstpi) + 1;
    jat = max(1, min(jat, h.jmax - 1));
    const int j0 = jat - 1, i0 = hz.i0;
    const double* base = h.node + (size_t)DS_HELM_REC * ((size_t)i0 + (size_t)h.imax * j0);
    Corners q{base, base + DS_HELM_REC, base + (size_t)DS_HELM_REC * h.imax,
              base + (size_t)DS_HELM_REC * h.imax + DS_HELM_REC};

    const double tj = __ldg(h.t + j0), tj1 = __ldg(h.t + j0 + 1);
    // helm_alloc.f:301-326 (same operations as the *_sav arrays)
    const Sd dt = S(tj1) - S(tj);
    const Sd dt2 = dt * dt;
    const Sd dti = S(__ddiv_rn(1.0, dt.v));
    const Sd dt2i = S(__ddiv_rn(1.0, dt2.v));

    Sd xt = (S(temp) - S(tj)) * dti;
    if (xt.v < 0.0) xt.v = 0.0;
    const Sd mxt = S(1.0) - xt;
    const W6& sd = hz.sd;
    const W6& dsd = hz.dsd;
    const W4& cd = hz.cd;

    // Two passes over the cell's 36 table values: the sums that share the density-derivative weights (dF/dd,
    // d2F/dd dT), then the ones that share the plain density weights (dF/dT, d2F/dT2, F).  The node pointer is laundered
    // between the passes so that the table values are loaded again (L1 hits) instead of being kept in registers.
    const W6 st{psi0(xt), psi1(xt) * dt, psi2(xt) * dt2, psi0(mxt), -psi1(mxt) * dt, psi2(mxt) * dt2};
    const W6 dst{dpsi0(xt) * dti, dpsi1(xt), dpsi2(xt) * dt, -dpsi0(mxt) * dti, dpsi1(mxt), -dpsi2(mxt) * dt};
    Sd free_, df_d, df_t, df_tt, df_dt;
    {
        const W6* const ts[2] = {&st, &dst};
        Sd r[2];
        h5n<2>(q, dsd, ts, r);
        df_d = r[0]; df_dt = r[1];
    }
    {
        const double* b2 = base;
        asm volatile("" : "+l"(b2));
        const Corners q2{b2, b2 + DS_HELM_REC, b2 + (size_t)DS_HELM_REC * h.imax, b2 + (size_t)DS_HELM_REC * h.imax + DS_HELM_REC};
        const W6 ddst{ddpsi0(xt) * dt2i, ddpsi1(xt) * dti, ddpsi2(xt), ddpsi0(mxt) * dt2i, -ddpsi1(mxt) * dti, ddpsi2(mxt)};
        const W6* const ts[3] = {&dst, &ddst, &st};
        Sd r[3];
        h5n<3>(q2, sd, ts, r);
        df_t = r[0]; df_tt = r[1]; free_ = r[2];
    }

    const W4 ct{xpsi0(xt), xpsi1(xt) * dt, xpsi0(mxt), -xpsi1(mxt) * dt};
    Sd dpepdd_in = h3(q, 9, ct, cd);
    if (dpepdd_in.v < 1.0e-30) dpepdd_in.v = 1.0e-30;
    const Sd etaele = h3(q, 13, ct, cd);

    const Sd yes = S(ye), x = S(din) * S(din);
    const Sd sele = -df_t * yes;
    o.pele = (x * df_d).v;
    o.dpept = (x * df_dt).v;
    o.dpepd = (dpepdd_in * yes).v;
    o.sele = sele.v;
    o.deept = (S(temp) * (-df_tt * yes)).v;
    o.eele = (yes * free_ + S(temp) * sele).v;
    o.etaele = etaele.v;
    if (!(o.sele > 0.0)) return -1;
    return 0;
}
