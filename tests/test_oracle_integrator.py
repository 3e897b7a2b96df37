"""The stiff integrator (RODAS driver + Steinebach's RODASP coefficients) is PARITY-UNPINNED by the
reference (MESA's isolve is not in the tree, NScool's own test has no light-curve golden).  These
tests pin what can be pinned without it: the coefficient set is a consistent 4th-order, stiffly
accurate Rosenbrock-Wanner method, and DOP853 reproduces known integrals."""
import numpy as np


def test_rodasp_nonlinear_accuracy(oracle):
    # y' = -y^2, y(0) = 1 -> 1/(1+t): error falls with the tolerance at >= 4th-order pace
    errs, steps = [], []
    for rtol in (1e-4, 1e-6, 1e-8):
        idid, y, c = oracle.rodasp_test(0, 1.0, rtol, rtol, 0.0, [1.0])
        assert idid == 1
        errs.append(abs(y[0] - 0.5)); steps.append(c[2])
    assert errs[0] < 1e-5 and errs[1] < 3e-8 and errs[2] < 5e-10
    order = np.log(errs[1] / errs[2]) / np.log(steps[2] / steps[1])
    assert order > 3.5, order


def test_rodasp_stiffly_accurate(oracle):
    # Prothero-Robinson, lambda = -1e6 (autonomous form): no order reduction, a handful of steps
    idid, y, c = oracle.rodasp_test(2, 1.0, 1e-6, 1e-6, 0.0, [1.0, 0.0])
    assert idid == 1 and c[2] < 20
    assert abs(y[0] - np.cos(1.0)) < 1e-8


def test_rodasp_tridiagonal_parabolic(oracle):
    # nonlinear heat equation on 64 cells with the banded (ml = mu = 1) LU: converges towards a tight-tolerance run
    u0 = np.linspace(1, 2, 64)
    ref = oracle.rodasp_test(1, 0.01, 1e-11, 1e-11, 0.0, u0)[1]
    prev = None
    for rtol in (1e-4, 1e-6):
        idid, y, c = oracle.rodasp_test(1, 0.01, rtol, rtol, 0.0, u0)
        err = np.max(np.abs(y - ref))
        assert idid == 1 and err < 30 * rtol
        if prev is not None:
            assert err < prev
        prev = err
