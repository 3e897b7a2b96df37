"""ctypes loader for the CPU oracle (oracle/liboracle.so) -- test infrastructure only.

The oracle is the checker: tests compare the CUDA product path against it. Nothing under
dstar_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")
HELM_PATH = os.path.join(ROOT, "dstar_b200", "data", "helm_table.bin")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def _ptr(a, typ=dp):
    if a is None:
        return typ()
    return a.ctypes.data_as(typ)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build():
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


class MicroTablesIn(C.Structure):
    _fields_ = [
        ("nz", C.c_int32), ("ncharged", C.c_int32),
        ("rho", dp), ("rho_bar", dp), ("P", dp), ("P_bar", dp), ("dm", dp),
        ("ionic", dp), ("ionic_bar", dp), ("Yion", dp), ("Yion_bar", dp),
        ("Zion", dp), ("Aion", dp), ("Tc_cell", dp), ("Tc_face", dp),
        ("eos_ctrl", C.c_double * 3), ("cond_ctrl", C.c_double * 6),
        ("nu_channels", C.c_int32 * 5), ("turn_on_shell_Urca", C.c_int32),
        ("shell_Urca_luminosity_coeff", C.c_double), ("lgP_shell_Urca", C.c_double),
        ("do_cells", C.c_int32), ("do_faces", C.c_int32), ("rho_bar1_override", C.c_double),
    ]


class EvolveIn(C.Structure):
    _fields_ = [
        ("nz", C.c_int32), ("n_tab", C.c_int32), ("tab_lnT", dp),
        ("tab_lnK", dp), ("tab_lnCp", dp), ("tab_lnGamma", dp), ("tab_lnEnu", dp),
        ("dm", dp), ("dm_bar", dp), ("area", dp), ("rho_bar", dp), ("ePhi", dp), ("e2Phi", dp),
        ("ePhi_bar", dp), ("e2Phi_bar", dp), ("P", dp),
        ("atm_nv", C.c_int32), ("atm_lgTb", dp), ("atm_lgTeff", dp), ("atm_lgflux", dp),
        ("heat", C.c_double * 9), ("turn_on_extra_heating", C.c_int32), ("enuc_override", dp),
        ("fix_core_temperature", C.c_int32), ("make_inner_boundary_insulating", C.c_int32),
        ("fix_atmosphere_temperature_when_accreting", C.c_int32),
        ("atmosphere_temperature_when_accreting", C.c_double),
        ("Mdot", C.c_double), ("epoch_start_time", C.c_double), ("epoch_duration", C.c_double),
        ("integration_tolerance", C.c_double), ("min_lg_T", C.c_double), ("max_lg_T", C.c_double),
        ("maximum_timestep", C.c_double), ("maximum_number_of_models", C.c_int32),
        ("starting_number_for_profile", C.c_int32), ("suppress_first_step_output", C.c_int32),
        ("reference_cost_quirks", C.c_int32),
    ]


class EvolveState(C.Structure):
    _fields_ = [
        ("lnT", dp), ("dt", C.c_double), ("tsec", C.c_double), ("model", C.c_int32),
        ("Teff", C.c_double), ("Lsurf", C.c_double), ("dlnLsdlnT", C.c_double),
        ("idid", C.c_int32), ("counters", C.c_int32 * 7),
    ]


class EvolveTrace(C.Structure):
    _fields_ = [
        ("capacity", C.c_int32), ("count", C.c_int32), ("model", ip),
        ("tsec", dp), ("dt", dp), ("Teff", dp), ("Lsurf", dp), ("Lnu", dp), ("Lnuc", dp),
        ("lnT_rows", dp),
    ]


class Oracle:
    """Thin numpy-facing wrapper over the oracle's C entry points."""

    def __init__(self, load_helm=True):
        self.lib = C.CDLL(build())
        L = self.lib
        L.dso_helm_load.argtypes = [C.c_char_p]
        L.dso_helm_eval.argtypes = [C.c_double] * 4 + [dp]
        L.dso_sf_load.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, dp, dp]
        L.dso_pcy97.argtypes = [C.c_double, C.c_double, C.c_int, dp, dp, dp]
        L.dso_rodasp_test.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, ip]
        self.helm_loaded = False
        if load_helm and os.path.exists(HELM_PATH):
            rc = L.dso_helm_load(HELM_PATH.encode())
            assert rc == 0, rc
            self.helm_loaded = True

    # ---- scalars / small helpers
    def constants(self):
        out = np.zeros(64)
        n = self.lib.dso_constants(_ptr(out), 64)
        return out[:n]

    def helm_eval(self, T, rho, abar, zbar):
        out = np.zeros(7)
        ierr = self.lib.dso_helm_eval(T, rho, abar, zbar, _ptr(out))
        return ierr, out

    def sf_load(self, typ, kF, Tc, kFmin=None, kFmax=None):
        kF = f64(kF); Tc = f64(Tc)
        self.lib.dso_sf_load(typ, len(kF), kF[0] if kFmin is None else kFmin,
                             kF[-1] if kFmax is None else kFmax, _ptr(kF), _ptr(Tc))

    def sf_scale(self, s):
        s = f64(s)
        self.lib.dso_sf_scale(_ptr(s))

    def sf_get_results(self, kFp, kFn):
        out = np.zeros(3)
        self.lib.dso_sf_get_results(C.c_double(kFp), C.c_double(kFn), _ptr(out))
        return out

    def moments(self, Z, A, abunds, mass_fracs=False, exclude_neutrons=False, renorm=False):
        Z = f64(Z); A = f64(A); ab = f64(abunds)
        n = len(Z)
        comp = np.zeros(11); xsum = C.c_double(0.0)
        idx = np.zeros(n, dtype=np.int32); Yion = np.zeros(n)
        nch = self.lib.dso_moments(n, _ptr(Z), _ptr(A), _ptr(ab), int(mass_fracs), int(exclude_neutrons),
                                   int(renorm), _ptr(comp), C.byref(xsum), _ptr(idx, ip), _ptr(Yion))
        return comp, xsum.value, nch, idx[:nch], Yion[:nch]

    def interp_pm(self, x, y):
        x = f64(x)
        f = np.zeros((len(x), 4))
        f[:, 0] = y
        self.lib.dso_interp_pm(_ptr(x), len(x), _ptr(f))
        return f

    def interp_value_and_slope(self, x, f, xv):
        x = f64(x); f = f64(f); xv = f64(np.atleast_1d(xv))
        val = np.zeros(len(xv)); slope = np.zeros(len(xv))
        self.lib.dso_interp_value_and_slope(_ptr(x), len(x), _ptr(f), len(xv), _ptr(xv), _ptr(val), _ptr(slope))
        return val, slope

    def pcy97(self, grav, Plight, lgTb):
        lgTb = f64(lgTb)
        a = np.zeros(len(lgTb)); b = np.zeros(len(lgTb))
        self.lib.dso_pcy97(grav, Plight, len(lgTb), _ptr(lgTb), _ptr(a), _ptr(b))
        return a, b

    def bc09_Teff(self, grav, Plight, Pb, lgTeff, tight=False):
        """do_get_bc09_Teff (bc09.f:45-95): returns ierr, lgTb, lgflux, info(n,6)."""
        lgTeff = f64(lgTeff); n = len(lgTeff)
        lgTb = np.zeros(n); lgflux = np.zeros(n); info = np.zeros((n, 6))
        self.lib.dso_bc09_Teff.argtypes = [C.c_double] * 3 + [C.c_int, dp, C.c_int, dp, dp, dp]
        rc = self.lib.dso_bc09_Teff(grav, Plight, Pb, n, _ptr(lgTeff), int(tight), _ptr(lgTb), _ptr(lgflux), _ptr(info))
        return rc, lgTb, lgflux, info

    def bc09_table(self, grav, Plight, Pb, nv=256, tight=False):
        """do_generate_atm_table('bc09') (dStar_atm_mod.f:46-106): returns ierr, lgTb(nv), lgTeff(nv,4), lgflux(nv,4)."""
        lgTb = np.zeros(nv); a = np.zeros((nv, 4)); b = np.zeros((nv, 4))
        self.lib.dso_bc09_table.argtypes = [C.c_double] * 3 + [C.c_int, C.c_int, dp, dp, dp]
        rc = self.lib.dso_bc09_table(grav, Plight, Pb, nv, int(tight), _ptr(lgTb), _ptr(a), _ptr(b))
        return rc, lgTb, a, b

    # ---- batched microphysics
    def eval_crust_eos(self, rho, T, ionic, Zion, Aion, Yion, Tcs, chi_in=None, eos_ctrl=None):
        rho = f64(rho); T = f64(T); ionic = f64(ionic); Yion = f64(Yion); Tcs = f64(Tcs)
        Zion = f64(Zion); Aion = f64(Aion)
        n = len(rho)
        nch = len(Zion)
        res = np.zeros((n, 15)); phase = np.zeros(n, dtype=np.int32); chi = np.zeros(n)
        chi_in = None if chi_in is None else f64(chi_in)
        ctrl = None if eos_ctrl is None else f64(eos_ctrl)
        ierr = self.lib.dso_eval_crust_eos(n, _ptr(rho), _ptr(T), _ptr(ionic), nch, _ptr(Zion), _ptr(Aion),
                                           _ptr(Yion), _ptr(Tcs), _ptr(chi_in), _ptr(ctrl), _ptr(res),
                                           _ptr(phase, ip), _ptr(chi))
        return ierr, res, phase, chi

    def crust_neutrino(self, rho, T, ionic, chi, Tcn, channels=(1, 1, 1, 1, 1)):
        rho = f64(rho); T = f64(T); ionic = f64(ionic); chi = f64(chi); Tcn = f64(Tcn)
        ch = np.asarray(channels, dtype=np.int32)
        eps = np.zeros((len(rho), 6))
        self.lib.dso_crust_neutrino(len(rho), _ptr(rho), _ptr(T), _ptr(ionic), _ptr(chi), _ptr(Tcn),
                                    _ptr(ch, ip), _ptr(eps))
        return eps

    def conductivity(self, rho, T, chi, Gamma, eta, mu_e, ionic, Tns, cond_ctrl):
        arrs = [f64(a) for a in (rho, T, chi, Gamma, eta, mu_e, ionic, Tns)]
        ctrl = f64(cond_ctrl)
        K = np.zeros((len(arrs[0]), 10))
        self.lib.dso_conductivity(len(arrs[0]), *[_ptr(a) for a in arrs], _ptr(ctrl), _ptr(K))
        return K

    def rodasp_test(self, problem, tend, rtol, atol, h0, y0):
        y = f64(y0).copy()
        cnt = np.zeros(7, dtype=np.int32)
        idid = self.lib.dso_rodasp_test(problem, tend, rtol, atol, h0, _ptr(y), _ptr(cnt, ip))
        return idid, y, cnt


def _micro_tables(self, nz, ncharged, rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, Yion, Yion_bar, Zion, Aion,
                  Tc_cell=None, Tc_face=None, eos_ctrl=(175.0, 140.0, float(np.float32(1e-9))),
                  cond_ctrl=(0, 0, 1, 1, 10.0, 9.0), nu_channels=(1, 1, 1, 1, 1), shell_Urca=(0, 1.0, 28.5),
                  do_cells=1, do_faces=1, rho_bar1_override=0.0):
    """do_setup_crust_transport for ONE star on the CPU oracle."""
    keep = [f64(x) for x in (rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, Yion, Yion_bar, Zion, Aion)]
    tc = None if Tc_cell is None else f64(Tc_cell)
    tf = None if Tc_face is None else f64(Tc_face)
    m = MicroTablesIn()
    m.nz = nz; m.ncharged = ncharged
    (m.rho, m.rho_bar, m.P, m.P_bar, m.dm, m.ionic, m.ionic_bar, m.Yion, m.Yion_bar, m.Zion, m.Aion) = \
        [_ptr(x) for x in keep]
    m.Tc_cell = _ptr(tc); m.Tc_face = _ptr(tf)
    m.eos_ctrl = (C.c_double * 3)(*eos_ctrl)
    m.cond_ctrl = (C.c_double * 6)(*cond_ctrl)
    m.nu_channels = (C.c_int32 * 5)(*nu_channels)
    m.turn_on_shell_Urca = int(shell_Urca[0]); m.shell_Urca_luminosity_coeff = shell_Urca[1]
    m.lgP_shell_Urca = shell_Urca[2]
    m.do_cells = do_cells; m.do_faces = do_faces; m.rho_bar1_override = rho_bar1_override
    n = 4 * 128 * nz
    lnT = np.zeros(128)
    out = [np.zeros(n) for _ in range(4)]
    rb1 = C.c_double(0.0)
    ierr = self.lib.dso_micro_tables(C.byref(m), _ptr(lnT), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]),
                                     C.byref(rb1))
    return {"ierr": ierr, "tab_lnT": lnT, "tab_lnCp": out[0], "tab_lnGamma": out[1], "tab_lnEnu": out[2],
            "tab_lnK": out[3], "rho_bar1": rb1.value}


def _evolve_epoch(self, ein_kwargs, lnT, dt, trace_cap=0):
    """do_integrate_crust for ONE star, ONE epoch.  ein_kwargs: dict of EvolveIn fields (arrays as numpy)."""
    e = EvolveIn()
    keep = []
    for k, v in ein_kwargs.items():
        if isinstance(v, np.ndarray):
            a = f64(v); keep.append(a); setattr(e, k, _ptr(a))
        elif k == "heat":
            e.heat = (C.c_double * 9)(*v)
        else:
            setattr(e, k, v)
    st = EvolveState()
    y = f64(lnT).copy()
    st.lnT = _ptr(y); st.dt = dt; st.tsec = 0.0; st.model = 0
    tr = None
    arrs = {}
    if trace_cap > 0:
        tr = EvolveTrace()
        tr.capacity = trace_cap; tr.count = 0
        arrs["model"] = np.zeros(trace_cap, dtype=np.int32)
        tr.model = _ptr(arrs["model"], ip)
        for nm in ("tsec", "dt", "Teff", "Lsurf", "Lnu", "Lnuc"):
            arrs[nm] = np.zeros(trace_cap); setattr(tr, nm, _ptr(arrs[nm]))
        arrs["lnT_rows"] = np.zeros((trace_cap, int(ein_kwargs["nz"])))
        tr.lnT_rows = _ptr(arrs["lnT_rows"])
    rc = self.lib.dso_evolve_epoch(C.byref(e), C.byref(st), C.byref(tr) if tr is not None else None)
    out = {"rc": rc, "lnT": y, "dt": st.dt, "tsec": st.tsec, "model": st.model, "Teff": st.Teff, "Lsurf": st.Lsurf,
           "idid": st.idid, "counters": np.array(list(st.counters))}
    if tr is not None:
        n = tr.count
        out["trace"] = {k: v[:n] for k, v in arrs.items()}
    return out


def _fill_evolve_in(ein_kwargs):
    e = EvolveIn()
    keep = []
    for k, v in ein_kwargs.items():
        if isinstance(v, np.ndarray):
            a = f64(v); keep.append(a); setattr(e, k, _ptr(a))
        elif k == "heat":
            e.heat = (C.c_double * 9)(*v)
        else:
            setattr(e, k, v)
    return e, keep


class RhsEvaluator:
    """get_derivatives / get_jacobian of ONE star and epoch at arbitrary ln T (function-level tests, independent
    integrations).  Keeps the filled EvolveIn alive, so repeated calls cost one C call each."""

    def __init__(self, oracle, ein_kwargs):
        self.lib = oracle.lib
        self.e, self._keep = _fill_evolve_in(ein_kwargs)
        self.n = int(ein_kwargs["nz"])

    def __call__(self, lnT):
        y = f64(lnT)
        n = self.n
        f = np.zeros(n); lo = np.zeros(n); di = np.zeros(n); up = np.zeros(n); surf = np.zeros(3)
        rc = self.lib.dso_evolve_rhs_jac(C.byref(self.e), _ptr(y), _ptr(f), _ptr(lo), _ptr(di), _ptr(up), _ptr(surf))
        assert rc == 0
        return {"f": f, "lo": lo, "di": di, "up": up, "Lsurf": surf[0], "dlnLsdlnT": surf[1], "Teff": surf[2]}


Oracle.micro_tables = _micro_tables
Oracle.evolve_epoch = _evolve_epoch
