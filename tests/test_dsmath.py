"""dsmath (deterministic elementary functions, the analogue of MESA's math_lib/crlibm layer):
accuracy of the CPU copy against numpy/libm on the argument ranges the hot path uses, and -- on
the GPU -- bit-for-bit equality of the device implementation with the CPU copy."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import _ptr, f64

FN = {"exp": 0, "log": 1, "log10": 2, "pow": 3, "sin": 4, "cos": 5, "sinh": 6, "cosh": 7, "atan": 8}


def cpu_eval(oracle, name, x, y=None):
    x = f64(x); y = f64(np.zeros_like(x) if y is None else y)
    out = np.zeros_like(x)
    oracle.lib.dso_dsmath(FN[name], len(x), _ptr(x), _ptr(y), _ptr(out))
    return out


def ulp_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return np.max(np.abs(got - ref) / np.spacing(np.abs(ref)))


def samples(rng):
    return {
        "exp": (rng.uniform(-700, 700, 20000), None, np.exp, 2.5),
        "log": (10.0 ** rng.uniform(-300, 300, 20000), None, np.log, 2.5),
        "log10": (10.0 ** rng.uniform(-30, 40, 20000), None, np.log10, 3.0),
        "sin": (rng.uniform(-60, 60, 20000), None, np.sin, None),
        "cos": (rng.uniform(-60, 60, 20000), None, np.cos, None),
        "sinh": (rng.uniform(0.5, 250, 20000), None, np.sinh, 4.0),
        "cosh": (rng.uniform(-250, 250, 20000), None, np.cosh, 4.0),
        "atan": (10.0 ** rng.uniform(-8, 8, 20000) * rng.choice([-1, 1], 20000), None, np.arctan, 6.0),
    }


def test_dsmath_accuracy(oracle):
    rng = np.random.default_rng(11)
    for name, (x, y, ref, tol) in samples(rng).items():
        got = cpu_eval(oracle, name, x, y)
        if tol is None:   # sin/cos: absolute accuracy (relative accuracy is lost near zeros by design)
            assert np.max(np.abs(got - ref(x))) < 4e-16, name
        else:
            assert ulp_err(got, ref(x)) < tol, (name, ulp_err(got, ref(x)))
    # pow: exact-root fast paths are correctly rounded-ish; general path carries |y ln x| ulp
    x = 10.0 ** rng.uniform(-40, 40, 20000)
    for yv in (0.5, 1.5, 2.5, 3.5, 0.25):
        assert ulp_err(cpu_eval(oracle, "pow", x, np.full_like(x, yv)), x ** yv) < 2.5
    for yv in (1 / 3, 2 / 3, 5 / 3, 0.617, -0.3, 8 / 3):
        got = cpu_eval(oracle, "pow", x, np.full_like(x, yv))
        assert np.max(np.abs(got / x ** yv - 1.0)) < 6e-14, yv
    # special values
    assert cpu_eval(oracle, "exp", [-1000.0, 0.0])[0] == 0.0 and cpu_eval(oracle, "exp", [0.0])[0] == 1.0
    assert cpu_eval(oracle, "log", [1.0])[0] == 0.0 and np.isinf(cpu_eval(oracle, "log", [0.0])[0])
    assert cpu_eval(oracle, "pow", [0.0, 5.0], [2.0, 0.0]).tolist() == [0.0, 1.0]
    # the branch-free forms: every special operand still gives the libm answer
    sp = np.array([np.inf, -np.inf, np.nan, 1e300, -1e300, 709.0, 709.7827, 709.79, 710.0, 800.0, -708.0, -740.0, -745.0,
                   -745.2, -746.0, -800.0, 5e-324, -5e-324])
    with np.errstate(all="ignore"):
        want = np.exp(sp)
    got = cpu_eval(oracle, "exp", sp)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    big = ok & np.isfinite(want) & (want > 1e-300)
    assert np.all(np.abs(got[big] / want[big] - 1.0) < 1e-15)
    assert np.array_equal(np.isinf(got[ok]), np.isinf(want[ok]))
    assert np.all(np.abs(got[ok & (want < 1e-300)] - want[ok & (want < 1e-300)]) <= 5e-324)   # subnormal range: 1 ulp
    sp = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0, -1.0, -1e-320, 5e-324, 1e-320, 2.2250738585072014e-308, 1.0,
                   1.7976931348623157e308, 1.0 + 2.0 ** -52, 1.0 - 2.0 ** -53, 1.4142135623730951, 0.7071067811865476])
    with np.errstate(all="ignore"):
        want = np.log(sp)
    got = cpu_eval(oracle, "log", sp)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok & np.isinf(want)], want[ok & np.isinf(want)])
    fin = ok & np.isfinite(want)
    assert ulp_err(got[fin & (want != 0)], want[fin & (want != 0)]) < 2.5 and got[10] == 0.0
    got = cpu_eval(oracle, "pow", [0.0, 0.0, 0.0], [1 / 3, -0.3, 2.066])
    assert got[0] == 0.0 and np.isinf(got[1]) and got[2] == 0.0


@pytest.mark.gpu
def test_dsmath_bits(oracle):
    """The device implementation returns exactly the bits of the CPU copy."""
    import dstar_b200 as ds
    from dstar_b200._capi import check, p
    ctx = ds.Context(0, gaps=None)
    rng = np.random.default_rng(12)
    try:
        cases = {k: (v[0], v[1]) for k, v in samples(rng).items()}
        xx = 10.0 ** rng.uniform(-40, 40, 40000)
        specials = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0, -1.0, 5e-324, 1e-320, 2.2250738585072014e-308, 1e300,
                             -1e300, 709.7827, 709.79, 710.0, 800.0, -745.0, -745.2, -746.0, -800.0, 1.0, 1.0 + 2.0 ** -52])
        for k in ("exp", "log", "log10"):
            cases[k] = (np.concatenate([cases[k][0], specials]), None)
        cases["pow"] = (xx, rng.choice([0.5, 1.5, 2.5, 0.25, 1 / 3, 2 / 3, 5 / 3, 0.617, -0.3, 8 / 3, 0.475, -2.066], 40000))
        for name, (x, y) in cases.items():
            x = f64(x); y = f64(np.zeros_like(x) if y is None else y)
            out = np.zeros_like(x)
            check(ctx.L.dstar_dsmath_eval(ctx.h, FN[name], len(x), p(x), p(y), p(out)), "dsmath_eval")
            ref = cpu_eval(oracle, name, x, y)
            nan = np.isnan(ref)          # NaN payload/sign is not part of the specification
            assert np.array_equal(np.isnan(out), nan), name
            assert np.array_equal(out.view(np.uint64)[~nan], ref.view(np.uint64)[~nan]), name
    finally:
        ctx.close()


@pytest.mark.gpu
def test_shared_denominator_division_bits():
    """DsDen (dstar_b200/csrc/dconst.cuh): the branch-free division returns exactly the bits of the IEEE quotient
    x / y -- moderate operands (fast path), extreme exponents, subnormals, zeros, infinities and NaN (flag raised,
    plain division).  One stated exception: a zero numerator gives a zero whose SIGN may differ from IEEE's (-0 / b
    with b > 0 comes out +0; see the comment in operator/)."""
    import dstar_b200 as ds
    from dstar_b200._capi import check, p
    ctx = ds.Context(0, gaps=None)
    rng = np.random.default_rng(77)
    try:
        def rnd(n, lo, hi):
            return rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(lo, hi, n) * rng.uniform(1.0, 2.0, n)
        n = 1 << 22
        sets = [(rnd(n, -30, 30), rnd(n, -30, 30)), (rnd(n, -300, 300), rnd(n, -300, 300)),
                (rnd(n, -3, 3), rnd(n, -3, 3))]
        # same significands, dense exponents near 1: half-way cases of the final rounding
        m = rng.integers(1 << 52, 1 << 53, n).astype(np.float64)
        sets.append((m * 2.0 ** -52, rng.integers(1 << 52, 1 << 53, n).astype(np.float64) * 2.0 ** -52))
        sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                       1e-300, 1e300, 3.0, 1 / 3, 2.0 ** -967, 2.0 ** 977, 2.0 ** -900, 2.0 ** 900])
        A, B = np.meshgrid(sp, sp)
        sets.append((A.ravel().copy(), B.ravel().copy()))
        for x, y in sets:
            x = f64(x); y = f64(y)
            q_plain = np.zeros_like(x); q_den = np.zeros_like(x)
            check(ctx.L.dstar_dsmath_eval(ctx.h, 9, len(x), p(x), p(y), p(q_plain)), "dsmath_eval")
            check(ctx.L.dstar_dsmath_eval(ctx.h, 10, len(x), p(x), p(y), p(q_den)), "dsmath_eval")
            with np.errstate(all="ignore"):
                host = x / y
            nan = np.isnan(host)
            assert np.array_equal(q_plain.view(np.uint64)[~nan], host.view(np.uint64)[~nan])
            assert np.array_equal(np.isnan(q_den), nan)
            zn = (x == 0.0) & ~nan          # zero numerators: a zero, sign not asserted
            ok = ~nan & ~zn
            assert np.array_equal(q_den.view(np.uint64)[ok], host.view(np.uint64)[ok])
            assert np.all(q_den[zn] == 0.0)
            # fast path alone: wherever the validity flag stays down the bits are the IEEE quotient's
            q_fast = np.zeros_like(x)
            check(ctx.L.dstar_dsmath_eval(ctx.h, 11, len(x), p(x), p(y), p(q_fast)), "dsmath_eval")
            took = ~np.isnan(q_fast)
            assert np.array_equal(q_fast.view(np.uint64)[took & ~zn], host.view(np.uint64)[took & ~zn])
            assert np.all(q_fast[took & zn] == 0.0)
            with np.errstate(all="ignore"):
                moderate = ((np.abs(np.log2(np.abs(y))) < 299) & (np.log2(np.abs(x)) > -959) & (np.log2(np.abs(host)) > -999)
                            & (np.log2(np.abs(host)) < 699) & ~nan)
            assert np.all(took[moderate])          # ... and the flag stays down for every moderate operand pair
    finally:
        ctx.close()


@pytest.mark.gpu
def test_branch_free_sqrt_bits():
    """ds_sqrt<true> (dconst.cuh) returns the bits of the IEEE square root for every finite x >= 2^-970 and for
    zero, and raises its flag (marked by NaN here) for everything else."""
    import dstar_b200 as ds
    from dstar_b200._capi import check, p
    ctx = ds.Context(0, gaps=None)
    rng = np.random.default_rng(78)
    try:
        n = 1 << 22
        sets = [10.0 ** rng.uniform(-30, 30, n), 10.0 ** rng.uniform(-291, 308, n), rng.uniform(1.0, 4.0, n),
                (rng.integers(1 << 52, 1 << 53, n).astype(np.float64)) * 2.0 ** -52,
                np.array([0.0, -0.0, 1.0, 4.0, 2.0 ** -970, 2.0 ** -971, 5e-324, 2.2250738585072014e-308,
                          1.7976931348623157e308, np.inf, -1.0, np.nan, 2.0, 3.0, 1e-300, 1e300])]
        for x in sets:
            x = f64(x); y = np.zeros_like(x)
            r_plain = np.zeros_like(x); r_fast = np.zeros_like(x)
            check(ctx.L.dstar_dsmath_eval(ctx.h, 12, len(x), p(x), p(y), p(r_plain)), "dsmath_eval")
            check(ctx.L.dstar_dsmath_eval(ctx.h, 13, len(x), p(x), p(y), p(r_fast)), "dsmath_eval")
            with np.errstate(all="ignore"):
                host = np.sqrt(x)
            ok = ~np.isnan(host)
            assert np.array_equal(r_plain.view(np.uint64)[ok], host.view(np.uint64)[ok])
            took = ~np.isnan(r_fast)
            assert np.array_equal(r_fast.view(np.uint64)[took], host.view(np.uint64)[took])
            safe = ((x >= 2.0 ** -970) & np.isfinite(x)) | (x == 0.0)
            assert np.array_equal(took, safe)
    finally:
        ctx.close()
