"""numpy-facing wrappers over the C ABI: Context (tables + point evaluators), Batch (HBM-resident
stars, seams A and B) and the NScool_lib mirror (alloc / setup / create_model / evolve_model).

Everything computes on the GPU through libdstar_b200.so; nothing here has a CPU path.
"""
import ctypes as C
import os

import numpy as np

from ._capi import DATA_DIR, DStarError, EvolveCtrl, MicroCtrl, check, dp, f64, i32, ip, lib, p

NTAB = 128


class Context:
    """dstar_ctx: device + read-only tables (replaces NScool_init / *_startup)."""

    def __init__(self, device=0, data_dir=DATA_DIR, gaps=("ns", "gc", "t72")):
        self.L = lib()
        self.h = C.c_void_p()
        check(self.L.dstar_ctx_create(int(device), C.byref(self.h)), "dstar_ctx_create")
        self.data_dir = data_dir
        helm = os.path.join(data_dir, "helm_table.bin")
        if check(self.L.dstar_ctx_load_helm_file(self.h, helm.encode()), "load_helm") != 0:
            raise DStarError("cannot read %s (run __graft_entry__.build())" % helm)
        if gaps is not None:
            self.load_gaps(*gaps)

    def load_gaps(self, p1, n1, n3):
        for typ, (name, code) in enumerate(((p1, "p1"), (n1, "n1"), (n3, "n3"))):
            path = os.path.join(self.data_dir, "Tc_data", "%s_%s.data" % (name, code))
            if check(self.L.dstar_ctx_load_gap_file(self.h, typ, path.encode()), "load_gap") != 0:
                raise DStarError("cannot read gap table " + path)

    def set_sf_scale(self, s):
        s = f64(s)
        check(self.L.dstar_ctx_set_sf_scale(self.h, p(s)), "set_sf_scale")

    def close(self):
        if self.h:
            self.L.dstar_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def device_info(self):
        sm = C.c_int32(); ma = C.c_int32(); mi = C.c_int32(); name = C.create_string_buffer(256)
        check(self.L.dstar_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), name), "device_info")
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "name": name.value.decode()}

    def fp64_peak_tflops(self):
        t = C.c_double()
        check(self.L.dstar_fp64_peak_tflops(self.h, C.byref(t)), "fp64_peak")
        return t.value

    # ---- batched point evaluators
    def eval_crust_eos(self, rho, T, ionic, Zion, Aion, Yion, Tcs, chi_in=None, eos_ctrl=None):
        rho = f64(rho); T = f64(T); ionic = f64(ionic); Yion = f64(Yion); Tcs = f64(Tcs)
        Zion = f64(Zion); Aion = f64(Aion)
        n = len(rho)
        res = np.zeros((n, 15)); phase = np.zeros(n, dtype=np.int32); chi = np.zeros(n)
        ierr = np.zeros(n, dtype=np.int32)
        chi_in = None if chi_in is None else f64(chi_in)
        ctrl = None if eos_ctrl is None else f64(eos_ctrl)
        rc = self.L.dstar_eval_crust_eos(self.h, n, p(rho), p(T), p(ionic), len(Zion), p(Zion), p(Aion), p(Yion),
                                         p(Tcs), p(chi_in), p(ctrl), p(res), p(phase, ip), p(chi), p(ierr, ip))
        check(rc, "dstar_eval_crust_eos")
        return ierr, res, phase, chi

    def crust_neutrino(self, rho, T, ionic, chi, Tcn, channels=(1, 1, 1, 1, 1)):
        rho = f64(rho); T = f64(T); ionic = f64(ionic); chi = f64(chi); Tcn = f64(Tcn)
        ch = i32(channels)
        eps = np.zeros((len(rho), 6))
        check(self.L.dstar_crust_neutrino(self.h, len(rho), p(rho), p(T), p(ionic), p(chi), p(Tcn), p(ch, ip),
                                          p(eps)), "dstar_crust_neutrino")
        return eps

    def conductivity(self, rho, T, chi, Gamma, eta, mu_e, ionic, Tns, cond_ctrl):
        a = [f64(x) for x in (rho, T, chi, Gamma, eta, mu_e, ionic, Tns)]
        ctrl = f64(cond_ctrl)
        K = np.zeros((len(a[0]), 10))
        check(self.L.dstar_conductivity(self.h, len(a[0]), *[p(x) for x in a], p(ctrl), p(K)), "dstar_conductivity")
        return K

    def sf_get_results(self, kFp, kFn):
        kFp = f64(np.atleast_1d(kFp)); kFn = f64(np.atleast_1d(kFn))
        Tc = np.zeros((len(kFn), 3))
        check(self.L.dstar_sf_get_results(self.h, len(kFn), p(kFp), p(kFn), p(Tc)), "dstar_sf_get_results")
        return Tc

    def crust_composition_info(self, lgP_tab, Ycoef, Ziso, Aiso, lgP):
        lgP_tab = f64(lgP_tab); Ycoef = f64(Ycoef); Ziso = f64(Ziso); Aiso = f64(Aiso); lgP = f64(lgP)
        nz = len(lgP)
        nch = int(np.sum((Ziso > 0) & (Aiso > 0)))
        ionic = np.zeros((nz, 11)); Yion = np.zeros((nz, nch)); n2 = C.c_int32()
        check(self.L.dstar_crust_composition_info(self.h, len(lgP_tab), len(Ziso), p(lgP_tab), p(Ycoef), p(Ziso),
                                                  p(Aiso), nz, p(lgP), p(ionic), p(Yion), C.byref(n2)),
              "dstar_crust_composition_info")
        return ionic, Yion, n2.value


def _atm_bc09_Teff(self, grav, Plight, Pb, lgTeff=None, with_info=False):
    """do_get_bc09_Teff (dStar_atm/private/bc09.f:45-95) for n_tab (grav, Plight, Pb) triples on the device.
    Returns lgTeff(nv), lgTb(n_tab, nv), lgflux(n_tab, nv), ierr(n_tab)[, info(n_tab, nv, 4)]."""
    grav = f64(np.atleast_1d(grav)); Plight = f64(np.atleast_1d(Plight)); Pb = f64(np.atleast_1d(Pb))
    if lgTeff is None:   # dStar_atm_def.f:4-8: 256 knots over lgTeff = 5.5 .. 7.0
        lgTeff = 5.5 + np.arange(256) * (7.0 - 5.5) / 255.0
    lgTeff = f64(lgTeff)
    nt, nv = len(grav), len(lgTeff)
    lgTb = np.zeros((nt, nv)); lgflux = np.zeros((nt, nv)); ierr = np.zeros(nt, dtype=np.int32)
    info = np.zeros((nt, nv, 4)) if with_info else None
    rc = self.L.dstar_atm_bc09_Teff(self.h, nt, p(grav), p(Plight), p(Pb), nv, p(lgTeff), p(lgTb), p(lgflux),
                                    p(ierr, ip), p(info))
    if rc <= -100:
        check(rc, "dstar_atm_bc09_Teff")
    return (lgTeff, lgTb, lgflux, ierr) + ((info,) if with_info else ())


Context.atm_bc09_Teff = _atm_bc09_Teff


def host_interp_pm(x, y):
    """MESA interp_pm on the host (setup utility of the product): returns the (n,4) coefficients."""
    x = f64(x)
    f = np.zeros((len(x), 4))
    f[:, 0] = y
    lib().dstar_host_interp_pm(p(x), len(x), p(f))
    return f


def abuntime_table(path, lgP_increment=0.005, abundance_threshold=0.01):
    """Preprocess an abuntime composition file as dStar_crust/preprocessor does (read, cull, reduce, extend with
    HZ90): returns (names, Z, A, lgP(nz), Y(nz, nnet)).  Host only."""
    import ctypes as C
    L = lib()
    nnet, nz = C.c_int32(), C.c_int32()
    bp = os.fsencode(path)
    rc = L.dstar_abuntime_table(bp, C.c_double(lgP_increment), C.c_double(abundance_threshold), C.byref(nnet),
                                C.byref(nz), None, None, None, None, None, 0, 0)
    if rc != 0:
        raise DStarError("abuntime_table: %s" % L.dstar_NScool_last_error().decode())
    names = C.create_string_buffer(6 * nnet.value)
    Z = np.zeros(nnet.value); A = np.zeros(nnet.value); lgP = np.zeros(nz.value); Y = np.zeros((nz.value, nnet.value))
    rc = L.dstar_abuntime_table(bp, C.c_double(lgP_increment), C.c_double(abundance_threshold), C.byref(nnet),
                                C.byref(nz), names, p(Z), p(A), p(lgP), p(Y), nnet.value, nz.value)
    if rc != 0:
        raise DStarError("abuntime_table: %s" % L.dstar_NScool_last_error().decode())
    nm = [names.raw[6 * k:6 * k + 6].split(b"\0")[0].decode() for k in range(nnet.value)]
    return nm, Z, A, lgP, Y


class Batch:
    """dstar_batch: stars flattened zone-major, resident in HBM."""

    def __init__(self, ctx, zone_offset, ncharged):
        self.ctx = ctx
        self.L = ctx.L
        self.zone_offset = i32(zone_offset)
        self.n_star = len(self.zone_offset) - 1
        self.total = int(self.zone_offset[-1])
        self.ncharged = int(ncharged)
        self.h = C.c_void_p()
        check(self.L.dstar_batch_create(ctx.h, self.n_star, p(self.zone_offset, ip), self.ncharged, C.byref(self.h)),
              "dstar_batch_create")
        self.n_epoch = 0

    def close(self):
        if self.h:
            self.L.dstar_batch_destroy(self.h)
            self.h = C.c_void_p()

    def upload_structure(self, rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, Yion, Yion_bar, Zion, Aion,
                         Tc_cell=None, Tc_face=None):
        a = [f64(x) for x in (rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, Yion, Yion_bar, Zion, Aion)]
        tc = None if Tc_cell is None else f64(Tc_cell)
        tf = None if Tc_face is None else f64(Tc_face)
        self._h2d_structure = sum(x.nbytes for x in a) + (tc.nbytes if tc is not None else 0) + \
            (tf.nbytes if tf is not None else 0)
        check(self.L.dstar_batch_upload_structure(self.h, *[p(x) for x in a], p(tc), p(tf)), "upload_structure")

    def set_table_sharing(self, which, owner):
        """which = 1 cell tables, 2 face tables; owner[i] = first star of star i's group, or None to end the sharing
        (dstar_batch_set_table_sharing)."""
        o = None if owner is None else i32(owner)
        check(self.L.dstar_batch_set_table_sharing(self.h, int(which), p(o, ip)), "dstar_batch_set_table_sharing")

    def micro_tables(self, ctrl=None, which=3):
        ctrl = ctrl or MicroCtrl.defaults()
        check(self.L.dstar_batch_micro_tables(self.h, C.byref(ctrl), int(which)), "dstar_batch_micro_tables")

    def download_tables(self, tables=True):
        n = 4 * NTAB * self.total
        lnT = np.zeros(NTAB)
        out = [np.zeros(n) if tables else None for _ in range(4)]
        rb1 = np.zeros(self.n_star); st = np.zeros(self.n_star, dtype=np.int32)
        check(self.L.dstar_batch_download_tables(self.h, p(lnT), *[p(x) for x in out], p(rb1), p(st, ip)),
              "download_tables")
        return {"tab_lnT": lnT, "tab_lnCp": out[0], "tab_lnGamma": out[1], "tab_lnEnu": out[2], "tab_lnK": out[3],
                "rho_bar1": rb1, "status": st}

    def upload_tables(self, tab_lnCp, tab_lnGamma, tab_lnEnu, tab_lnK):
        a = [f64(x) for x in (tab_lnCp, tab_lnGamma, tab_lnEnu, tab_lnK)]
        check(self.L.dstar_batch_upload_tables(self.h, None, *[p(x) for x in a]), "upload_tables")

    def upload_evolve(self, dm_bar, area, ePhi, e2Phi, ePhi_bar, e2Phi_bar, atm_lgTb, atm_lgTeff, atm_lgflux,
                      atm_index, heat, T_atm, lnT0, epoch_boundaries, epoch_Mdots, enuc_per_Mdot=None):
        geo = [f64(x) for x in (dm_bar, area, ePhi, e2Phi, ePhi_bar, e2Phi_bar)]
        atm_lgTb = f64(np.atleast_2d(atm_lgTb)); atm_lgTeff = f64(np.atleast_2d(atm_lgTeff))
        atm_lgflux = f64(np.atleast_2d(atm_lgflux))
        n_atm, nv = atm_lgTb.shape
        atm_index = i32(atm_index); heat = f64(heat); T_atm = f64(T_atm); lnT0 = f64(lnT0)
        eb = f64(epoch_boundaries); em = f64(epoch_Mdots)   # (n_epoch+1, n_star), (n_epoch, n_star)
        self.n_epoch = em.shape[0]
        en = None if enuc_per_Mdot is None else f64(enuc_per_Mdot)
        self._h2d_evolve = sum(x.nbytes for x in geo) + atm_lgTb.nbytes + atm_lgTeff.nbytes + atm_lgflux.nbytes + \
            heat.nbytes + T_atm.nbytes + lnT0.nbytes + eb.nbytes + em.nbytes
        check(self.L.dstar_batch_upload_evolve(self.h, *[p(x) for x in geo], n_atm, nv, p(atm_lgTb), p(atm_lgTeff),
                                               p(atm_lgflux), p(atm_index, ip), p(heat), p(T_atm), p(lnT0),
                                               self.n_epoch, p(eb), p(em), p(en)), "upload_evolve")

    def reset_state(self):
        check(self.L.dstar_batch_reset_state(self.h), "dstar_batch_reset_state")

    def evolve(self, ctrl=None, trace_cap=0):
        ctrl = ctrl or EvolveCtrl.defaults()
        self.trace_cap = trace_cap
        check(self.L.dstar_batch_evolve(self.h, C.byref(ctrl), int(trace_cap)), "dstar_batch_evolve")

    def set_evolve_mapping(self, mapping):
        """0 automatic, 1 CTA per star, 2 warp per star (dstar_batch_set_evolve_mapping)."""
        check(self.L.dstar_batch_set_evolve_mapping(self.h, int(mapping)), "dstar_batch_set_evolve_mapping")

    def set_evolve_occupancy(self, ctas_per_sm):
        check(self.L.dstar_batch_set_evolve_occupancy(self.h, int(ctas_per_sm)), "dstar_batch_set_evolve_occupancy")

    def eval_rhs(self, epoch, ctrl=None):
        """get_derivatives and the rows of get_jacobian at the batch's current ln T (dstar_batch_eval_rhs)."""
        ctrl = ctrl or EvolveCtrl.defaults()
        f = np.zeros(self.total); jac = np.zeros((self.total, 3)); surf = np.zeros((self.n_star, 3))
        check(self.L.dstar_batch_eval_rhs(self.h, C.byref(ctrl), int(epoch), p(f), p(jac), p(surf)), "dstar_batch_eval_rhs")
        return {"f": f, "lo": jac[:, 0].copy(), "di": jac[:, 1].copy(), "up": jac[:, 2].copy(), "Lsurf": surf[:, 0].copy(),
                "dlnLsdlnT": surf[:, 1].copy(), "Teff": surf[:, 2].copy()}

    def download_results(self, full=True):
        ne, ns = self.n_epoch, self.n_star
        t = np.zeros((ne, ns)); Te = np.zeros((ne, ns)); Qb = np.zeros((ne, ns))
        idid = np.zeros((ne, ns), dtype=np.int32); cnt = np.zeros((ns, 7), dtype=np.int32)
        lnT = np.zeros(self.total) if full else None
        Teff = np.zeros(ns); Ls = np.zeros(ns); dt = np.zeros(ns); ts = np.zeros(ns); model = np.zeros(ns, dtype=np.int32)
        check(self.L.dstar_batch_download_results(self.h, p(t), p(Te), p(Qb), p(idid, ip), p(cnt, ip), p(lnT), p(Teff),
                                                  p(Ls), p(dt), p(ts), p(model, ip)), "download_results")
        return {"t_monitor": t, "Teff_monitor": Te, "Qb_monitor": Qb, "idid": idid, "counters": cnt, "lnT": lnT,
                "Teff": Teff, "Lsurf": Ls, "dt": dt, "tsec": ts, "model": model}

    def download_trace(self):
        cap, ns = self.trace_cap, self.n_star
        cnt = np.zeros(ns, dtype=np.int32); model = np.zeros((cap, ns), dtype=np.int32)
        arr = [np.zeros((cap, ns)) for _ in range(7)]
        check(self.L.dstar_batch_download_trace(self.h, p(cnt, ip), p(model, ip), *[p(x) for x in arr]), "download_trace")
        names = ["tsec", "dt", "Teff", "Lsurf", "Lnu", "Lnuc", "Mdot"]
        out = {"count": cnt, "model": model}
        out.update(dict(zip(names, arr)))
        return out

    def kernel_ms(self, which):
        return self.L.dstar_batch_last_kernel_ms(self.h, int(which))

    def launch_count(self):
        """kernels launched by the last micro_tables / evolve call"""
        return int(self.L.dstar_batch_last_launch_count(self.h))

    def redo_counts(self):
        out = np.zeros(2, dtype=np.int32)
        check(self.L.dstar_batch_last_redo_counts(self.h, p(out, ip)), "last_redo_counts")
        return out


# ---------------------------------------------------------------------------- NScool_lib mirror
_initialised = False


def NScool_init(data_dir=DATA_DIR, device=0):
    """NScool_init(my_dStar_dir, ierr) -- NScool/public/NScool_lib.f:6."""
    global _initialised
    L = lib()
    L.dstar_NScool_last_error.restype = C.c_char_p
    L.dstar_NScool_last_kernel_ms.restype = C.c_double
    rc = L.dstar_NScool_init(data_dir.encode(), int(device))
    if rc != 0:
        raise DStarError("NScool_init failed (%d): %s / %s" % (rc, L.dstar_last_error().decode(),
                                                              L.dstar_NScool_last_error().decode()))
    _initialised = True


def NScool_shutdown():
    """NScool_shutdown -- NScool_lib.f:14."""
    global _initialised
    if _initialised:
        lib().dstar_NScool_shutdown()
    _initialised = False


def _nerr(rc, what):
    if rc < 0:
        L = lib()
        raise DStarError("%s: ierr = %d: %s | %s" % (what, rc, L.dstar_NScool_last_error().decode(),
                                                    L.dstar_last_error().decode()))
    return rc


class NScool:
    """One NScool_info slot (alloc_NScool / NScool_setup / NScool_create_model / NScool_evolve_model)."""

    def __init__(self, inlist=None, **controls):
        self.L = lib()
        self.id = _nerr(self.L.dstar_alloc_NScool(), "alloc_NScool")
        _nerr(self.L.dstar_NScool_setup(self.id, (inlist or "").encode()), "NScool_setup")
        self._hooks = None
        for k, v in controls.items():
            self.set_control(k, v)

    def set_control(self, name, value):
        if isinstance(value, bool):
            s = ".TRUE." if value else ".FALSE."
        elif isinstance(value, (list, tuple, np.ndarray)):
            s = ",".join(repr(float(x)) for x in value)
        elif isinstance(value, str):
            s = "'%s'" % value
        else:
            s = repr(value)
        _nerr(self.L.dstar_NScool_set_control(self.id, name.encode(), s.encode()), "set_control(%s)" % name)

    def set_hooks(self, other_set_Qimp=None, other_sf_get_results=None):
        QF = C.CFUNCTYPE(None, C.c_int, ip)
        SF = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_double, dp)
        q = QF(other_set_Qimp) if other_set_Qimp else None
        s = SF(other_sf_get_results) if other_sf_get_results else None
        self._hooks = (q, s)   # keep the trampolines alive
        self.L.dstar_NScool_set_hooks.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _nerr(self.L.dstar_NScool_set_hooks(self.id, C.cast(q, C.c_void_p) if q else None,
                                            C.cast(s, C.c_void_p) if s else None), "set_hooks")

    def use_example_hooks(self):
        """other_set_Qimp => alt_Qimp, other_sf_get_results => alt_sf (examples/custom_Tc_Qimp/src/run.f:50-51),
        the native versions shipped in the library."""
        L = self.L
        L.dstar_NScool_set_hooks.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _nerr(L.dstar_NScool_set_hooks(self.id, C.cast(L.dstar_example_alt_Qimp, C.c_void_p),
                                       C.cast(L.dstar_example_alt_sf, C.c_void_p)), "set_hooks")

    def create_model(self):
        _nerr(self.L.dstar_NScool_create_model(self.id), "NScool_create_model")

    def evolve_model(self):
        return _nerr(self.L.dstar_NScool_evolve_model(self.id), "NScool_evolve_model")

    def free(self):
        self.L.dstar_dealloc_NScool(self.id)

    # accessors
    def get_int(self, name):
        v = C.c_int32()
        _nerr(self.L.dstar_NScool_get_int(self.id, name.encode(), C.byref(v)), "get_int " + name)
        return v.value

    def get_real(self, name):
        v = C.c_double()
        _nerr(self.L.dstar_NScool_get_real(self.id, name.encode(), C.byref(v)), "get_real " + name)
        return v.value

    def array(self, name):
        n = _nerr(self.L.dstar_NScool_get_zone_array(self.id, name.encode(), None, 0), "array " + name)
        out = np.zeros(n)
        self.L.dstar_NScool_get_zone_array(self.id, name.encode(), p(out), n)
        return out

    def set_array(self, name, values):
        v = f64(values)
        _nerr(self.L.dstar_NScool_set_zone_array(self.id, name.encode(), p(v), v.size), "set_array " + name)

    def monitors(self):
        ne = self.get_int("number_epochs")
        t = np.zeros(ne); Te = np.zeros(ne); Qb = np.zeros(ne)
        _nerr(self.L.dstar_NScool_get_monitors(self.id, p(t), p(Te), p(Qb)), "get_monitors")
        return t, Te, Qb

    def counters(self):
        c = np.zeros(7, dtype=np.int32)
        _nerr(self.L.dstar_NScool_get_counters(self.id, p(c, ip)), "get_counters")
        return c

    def history(self):
        cap = 8192
        model = np.zeros(cap, dtype=np.int32)
        cols = [np.zeros(cap) for _ in range(6)]
        n = _nerr(self.L.dstar_NScool_get_history(self.id, cap, p(model, ip), *[p(x) for x in cols]), "get_history")
        names = ["time", "lg_Mdot", "lg_Teff", "lg_Lsurf", "lg_Lnu", "lg_Lnuc"]
        out = {"model": model[:n]}
        out.update({k: v[:n] for k, v in zip(names, cols)})
        return out

    def write_history(self, path):
        """Returns 0, or 1 when the file was truncated at the trace capacity."""
        return _nerr(self.L.dstar_NScool_write_history(self.id, path.encode()), "write_history")

    def write_terminal(self, path=None):
        return _nerr(self.L.dstar_NScool_write_terminal(self.id, None if path is None else path.encode()), "write_terminal")

    def terminal(self):
        cap = 8192
        model = np.zeros(cap, dtype=np.int32)
        cols = [np.zeros(cap) for _ in range(6)]
        n = _nerr(self.L.dstar_NScool_get_terminal(self.id, cap, p(model, ip), *[p(x) for x in cols]), "get_terminal")
        names = ["tsec", "dt", "lgTmax", "lgP_Tmax", "lgTmin", "lgP_Tmin"]
        out = {"model": model[:n]}
        out.update({k: v[:n] for k, v in zip(names, cols)})
        return out

    PROFILE_COLUMNS = ("zone", "mass", "dm", "area", "gravity", "eLambda", "ePhi", "temperature", "luminosity", "pressure",
                       "density", "zbar", "abar", "Xn", "Qimp", "Gamma", "cp", "eps_nu", "eps_nuc", "K")

    def profile_count(self):
        """(profiles kept, profiles the run produced): mod(model, write_interval_for_profile) == 0."""
        kept, made = C.c_int32(), C.c_int32()
        _nerr(self.L.dstar_NScool_profile_count(self.id, C.byref(kept), C.byref(made)), "profile_count")
        return kept.value, made.value

    def profile(self, index):
        """One captured profile: dict with the header scalars and the 20 zone columns of NScool_profile.f."""
        nz = self.get_int("nz")
        model = C.c_int32(); hdr = np.zeros(8); cols = np.zeros((nz, 20))
        _nerr(self.L.dstar_NScool_get_profile(self.id, index, C.byref(model), p(hdr), p(cols)), "get_profile")
        out = {"model": model.value}
        out.update(dict(zip(("time", "core_mass", "core_radius", "accretion_rate", "heating_lum", "cooling_lum",
                             "total_mass", "total_radius"), hdr)))
        out.update({k: cols[:, j].copy() for j, k in enumerate(self.PROFILE_COLUMNS)})
        return out

    def write_profiles(self, directory):
        os.makedirs(directory, exist_ok=True)
        _nerr(self.L.dstar_NScool_write_profiles(self.id, os.fsencode(directory)), "write_profiles")


def create_models(stars):
    ids = i32([s.id for s in stars])
    _nerr(lib().dstar_NScool_create_models(len(ids), p(ids, ip)), "NScool_create_models")


def table_sets():
    """(distinct cell-table sets, distinct face-table sets) the last create_models evaluated."""
    a, b = C.c_int32(), C.c_int32()
    lib().dstar_NScool_table_sets(C.byref(a), C.byref(b))
    return a.value, b.value


def evolve_models(stars):
    ids = i32([s.id for s in stars])
    return _nerr(lib().dstar_NScool_evolve_models(len(ids), p(ids, ip)), "NScool_evolve_models")


def run_batch(stars, n_gpu=1, devices=None):
    """dstar_NScool_run_batch: create + evolve `stars` on n_gpu devices of this node, one host thread per device,
    contiguous blocks of stars per device.  Returns (t_monitor, Teff_monitor, Qb_monitor) as (number_epochs, n)
    arrays and the per-star status (0 or the most negative isolve code)."""
    ids = i32([s.id for s in stars])
    n = len(ids)
    st = np.zeros(n, dtype=np.int32)
    dv = i32(devices) if devices is not None else None
    L = lib()
    # first call: no monitor buffers (number_epochs is known only after create_model has read the epochs)
    _nerr(L.dstar_NScool_run_batch(int(n_gpu), p(dv, ip), n, p(ids, ip), None, None, None, p(st, ip)), "NScool_run_batch")
    mons = [s.monitors() for s in stars]
    ne = max(len(m[0]) for m in mons) if mons else 0
    out = [np.zeros((ne, n)) for _ in range(3)]
    for i, m in enumerate(mons):
        for q in range(3):
            out[q][:len(m[q]), i] = m[q]
    return out[0], out[1], out[2], st


def last_kernel_ms(which):
    return lib().dstar_NScool_last_kernel_ms(int(which))
