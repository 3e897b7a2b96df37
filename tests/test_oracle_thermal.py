"""Pinning what can be pinned of seam B without MESA (isolve/rodasp is absent from the reference tree and NScool holds
no light-curve golden):

* the analytic Jacobian (get_jacobian, NScool/private/NScool_evolve.f:285-359) against central finite differences of
  the right-hand side (get_derivatives, :237-283) -- what the reference's own get_num_jacobian (:361-431) is for;
* the oracle's RODASP driver against an INDEPENDENT stiff integrator (scipy's Radau IIA, order 5) on the actual crust
  problem: the same right-hand side integrated to 1e-10 gives the converged T_eff(t); the RODASP run at a tight
  tolerance must agree with it, and the run at the reference's tolerance (1e-4) shows how far from converged the
  reference's own light curve is -- which bounds what "1e-6 against the Fortran path" can mean.

Inputs: tests/golden/basic_run_small.npz (a 21-zone basic_run star with its device-built coefficient tables)."""
import os

import numpy as np
import pytest

from oracle_lib import RhsEvaluator
from parity_helpers import evolve_in_kwargs, oracle_evolve_star
from test_regression_fixture import TABS, FrozenStar

HERE = os.path.dirname(os.path.abspath(__file__))
HEAT = [26.86, 30.13, 0.3, 30.42, 31.20, 1.5, 27.5, 28.5, 0.1]


@pytest.fixture(scope="module")
def star():
    fx = dict(np.load(os.path.join(HERE, "golden", "basic_run_small.npz")))
    tabs = {k: fx["gpu_" + k] for k in TABS}
    tabs["tab_lnT"] = fx["tab_lnT"]
    return FrozenStar(fx), tabs, fx


def _fd_rows(ev, y, h):
    """Central differences, column by column as get_num_jacobian (:398-430) does; returned as rows."""
    n = len(y)
    lo = np.zeros(n); di = np.zeros(n); up = np.zeros(n)
    for j in range(n):
        yp = y.copy(); ym = y.copy()
        yp[j] += h; ym[j] -= h
        d = (ev(yp)["f"] - ev(ym)["f"]) / (2.0 * h)
        di[j] = d[j]
        if j > 0:
            up[j - 1] = d[j - 1]     # d f_{j-1} / d y_j
        if j < n - 1:
            lo[j + 1] = d[j + 1]     # d f_{j+1} / d y_j
    return lo, di, up


@pytest.mark.parametrize("epoch,insulating", [(0, False), (1, False), (1, True)])
def test_analytic_jacobian_matches_finite_differences(oracle, star, epoch, insulating):
    s, tabs, fx = star
    ctrl = dict(turn_on_extra_heating=1)
    if insulating:
        ctrl.update(fix_core_temperature=0, make_inner_boundary_insulating=1)
    kw = evolve_in_kwargs(s, epoch, tabs=tabs, ctrl=ctrl, heat=HEAT, T_atm=2.4e8)
    ev = RhsEvaluator(oracle, kw)
    n = kw["nz"]
    # a realistic non-isothermal state: the profile at the end of the stored run, and the initial one
    for y0 in (fx["gpu_lnT"].astype(float), fx["lnT0"].astype(float)):
        y = y0.copy()
        fixed = epoch == 0    # accreting with a fixed atmosphere temperature: y(1) is pinned (:69-71)
        if fixed:
            y[0] = np.log(2.4e8)
        a = ev(y)
        lo, di, up = _fd_rows(ev, y, 1.0e-6)
        if fixed:   # the reference zeroes row 1 and column 1 (:354-357): y(1) is not an unknown
            lo[1] = 0.0; di[0] = 0.0; up[0] = 0.0
            assert a["lo"][1] == 0.0 and a["di"][0] == 0.0 and a["up"][0] == 0.0
        # "dfdy(3,n-1) = 0.0 ! for the isothermal core" (:352) is applied in the insulating case too: reference behaviour
        lo[n - 1] = 0.0
        assert a["lo"][n - 1] == 0.0
        scale = np.maximum.reduce([np.abs(di), np.abs(lo), np.abs(up)]) + 1e-300
        for name, fd in (("lo", lo), ("di", di), ("up", up)):
            dev = np.max(np.abs(a[name] - fd) / scale)
            assert dev < 1.0e-6, (name, epoch, insulating, dev)
        if not insulating:   # fixed core temperature: the last equation is y' = 0 (:276-280)
            assert a["f"][n - 1] == 0.0 and a["di"][n - 1] == 0.0 and a["up"][n - 2] != 0.0


def _radau_history(oracle, s, tabs, lnT0, rtol, atol):
    """The two epochs of the fixture star with scipy's Radau on the oracle's right-hand side and Jacobian."""
    from scipy.integrate import solve_ivp
    y = np.array(lnT0, dtype=float)
    eb = s.array("epoch_boundaries"); em = s.array("epoch_Mdots")
    Teff = []
    for ie in range(len(em)):
        kw = evolve_in_kwargs(s, ie, tabs=tabs, ctrl=dict(turn_on_extra_heating=1), heat=HEAT, T_atm=2.4e8)
        ev = RhsEvaluator(oracle, kw)
        n = kw["nz"]
        if em[ie] > 0.0:
            y[0] = np.log(2.4e8)

        def fun(t, yy):
            return ev(yy)["f"]

        def jac(t, yy):
            a = ev(yy)
            J = np.diag(a["di"]) + np.diag(a["up"][:-1], 1) + np.diag(a["lo"][1:], -1)
            return J

        r = solve_ivp(fun, (0.0, float(eb[ie + 1] - eb[ie])), y, method="Radau", jac=jac, rtol=rtol, atol=atol)
        assert r.success, r.message
        y = r.y[:, -1].copy()
        Teff.append(ev(y)["Teff"] * s.array("ePhi")[0])
    return np.array(Teff), y


def test_rodasp_light_curve_against_independent_integrator(oracle, star):
    s, tabs, fx = star
    lnT0 = fx["lnT0"]
    Teff_ref, y_ref = _radau_history(oracle, s, tabs, lnT0, rtol=1.0e-10, atol=1.0e-12)
    # (1) RODASP as implemented here, tight tolerance: agrees with the independent converged solution
    tight = oracle_evolve_star(oracle, s, lnT0, tabs=dict(tabs), heat=HEAT, T_atm=2.4e8,
                               ctrl=dict(turn_on_extra_heating=1, integration_tolerance=1.0e-9))
    d_tight = np.max(np.abs(tight["Teff_monitor"] / Teff_ref - 1.0))
    d_tight_y = np.max(np.abs(tight["lnT"] - y_ref))
    # (2) the reference's own setting (integration_tolerance = 1e-4, controls.defaults:18; atol = tol*min(lnT), :76)
    loose = oracle_evolve_star(oracle, s, lnT0, tabs=dict(tabs), heat=HEAT, T_atm=2.4e8, ctrl=dict(turn_on_extra_heating=1))
    d_loose = np.max(np.abs(loose["Teff_monitor"] / Teff_ref - 1.0))
    print("T_eff(t) vs Radau(1e-10): RODASP tol 1e-9 -> %.2e (ln T %.2e); RODASP tol 1e-4 (reference setting) -> %.2e"
          % (d_tight, d_tight_y, d_loose))
    assert d_tight < 1.0e-6 and d_tight_y < 1.0e-6
    # the reference's tolerance leaves the light curve converged to about 1e-3..1e-4 only: agreement to 1e-6 with the
    # Fortran path therefore means "same step sequence", not "same converged answer"
    assert d_loose < 5.0e-3
