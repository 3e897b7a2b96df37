"""numpy model of the evolve kernel's partitioned tridiagonal solve (dstar_b200/csrc/kernels_evolve.cuh: part_factor /
part_solve): chunks of 32 rows reduced by cyclic reduction (the shuffles become array shifts), spikes, the tridiagonal
interface system in (f(0), l(0), f(1), l(1), ...) solved by the same reduction.  Checked against dense solves for the
chunk layouts the kernel meets (partial last chunk, a single chunk, 16 chunks) and for weakly dominant diffusion-like
rows, where it must be as accurate as the whole-system cyclic reduction it replaced.  CPU only: this pins the ALGEBRA
the kernel implements; the kernel itself is compared with the oracle in the -m gpu tests."""
import numpy as np


def _up(x, s):     # __shfl_up_sync: lane i reads lane i - s, its own value below lane s
    y = x.copy(); y[:, s:] = x[:, :-s]; return y


def _down(x, s):   # __shfl_down_sync
    y = x.copy(); y[:, :-s] = x[:, s:]; return y


def _cr_factor(a, b, c, nlev):
    al, ga = [], []
    lane = np.arange(a.shape[1])[None, :]
    for lev in range(nlev):
        s = 1 << lev
        binv = 1.0 / b
        l = np.where(lane - s >= 0, -a * _up(binv, s), 0.0)
        g = np.where(lane + s < a.shape[1], -c * _down(binv, s), 0.0)
        b = b + l * _up(c, s) + g * _down(a, s)
        a, c = l * _up(a, s), g * _down(c, s)
        al.append(l); ga.append(g)
    return al, ga, 1.0 / b


def _cr_solve(f, d):
    al, ga, binv = f
    for lev in range(len(al)):
        s = 1 << lev
        d = d + al[lev] * _up(d, s) + ga[lev] * _down(d, s)
    return d * binv


def partitioned_solve(n, a, b, c, d, nw):
    N = 32 * nw
    A = np.zeros(N); B = np.ones(N); C = np.zeros(N); D = np.zeros(N)      # rows past n: identity
    A[:n], B[:n], C[:n], D[:n] = a, b, c, d
    A, B, C, D = (x.reshape(nw, 32) for x in (A, B, C, D))
    a_first, c_last = A[:, 0].copy(), C[:, 31].copy()
    Al, Cl = A.copy(), C.copy()
    Al[:, 0] = 0.0; Cl[:, 31] = 0.0
    f = _cr_factor(Al, B.copy(), Cl, 5)
    e0 = np.zeros((nw, 32)); e0[:, 0] = a_first
    e1 = np.zeros((nw, 32)); e1[:, 31] = c_last
    v, u = _cr_solve(f, e0), _cr_solve(f, e1)
    vf, vl, uf, ul = v[:, 0], v[:, 31], u[:, 0], u[:, 31]
    ru = np.where(ul != 0, uf / np.where(ul != 0, ul, 1.0), 0.0)
    rv = np.where(vf != 0, vl / np.where(vf != 0, vf, 1.0), 0.0)
    RA, RB, RC = np.zeros(32), np.ones(32), np.zeros(32)
    RA[0:2 * nw:2] = vf - vl * ru; RC[0:2 * nw:2] = -ru
    RA[1:2 * nw:2] = -rv;          RC[1:2 * nw:2] = ul - uf * rv
    rf = _cr_factor(RA[None, :], RB[None, :], RC[None, :], 4 if nw <= 8 else 5)
    y = _cr_solve(f, D)
    yf, yl = y[:, 0], y[:, 31]
    RD = np.zeros(32)
    RD[0:2 * nw:2] = yf - ru * yl; RD[1:2 * nw:2] = yl - rv * yf
    xr = _cr_solve(rf, RD[None, :])[0]
    w = np.arange(nw)
    lm = xr[np.maximum(2 * w - 1, 0)]          # chunk 0 has v = 0, the last chunk u = 0: their lm / fp do not enter
    fp = xr[np.minimum(2 * w + 2, 31)]
    return (y - lm[:, None] * v - fp[:, None] * u).reshape(-1)[:n]


def whole_system_cr(n, a, b, c, d):
    a, b, c, d = a.copy(), b.copy(), c.copy(), d.copy()
    s = 1
    while s < n:
        binv = 1.0 / b
        al = np.zeros(n); ga = np.zeros(n)
        al[s:] = -a[s:] * binv[:-s]; ga[:-s] = -c[:-s] * binv[s:]
        nb, na, nc, nd = b.copy(), np.zeros(n), np.zeros(n), d.copy()
        nb[s:] += al[s:] * c[:-s]; na[s:] = al[s:] * a[:-s]; nd[s:] += al[s:] * d[:-s]
        nb[:-s] += ga[:-s] * a[s:]; nc[:-s] = ga[:-s] * c[s:]; nd[:-s] += ga[:-s] * d[s:]
        a, b, c, d = na, nb, nc, nd
        s *= 2
    return d / b


def _dense(a, b, c):
    return np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)


def test_partitioned_solve_matches_dense_solve():
    rng = np.random.default_rng(20261020)
    for n, nw in [(253, 8), (256, 8), (3, 1), (31, 1), (33, 2), (64, 2), (65, 3), (97, 4), (225, 8), (271, 9), (500, 16), (512, 16)]:
        worst = 0.0
        for _ in range(10):
            a = -rng.uniform(0.1, 1, n); c = -rng.uniform(0.1, 1, n); a[0] = 0; c[-1] = 0
            b = np.abs(a) + np.abs(c) + rng.uniform(0.01, 2, n) * 10 ** rng.uniform(-3, 3)
            d = rng.normal(size=n)
            xe = np.linalg.solve(_dense(a, b, c), d)
            worst = max(worst, np.max(np.abs(partitioned_solve(n, a, b, c, d, nw) - xe)) / np.max(np.abs(xe)))
        assert worst < 1e-12, (n, nw, worst)


def test_partitioned_solve_on_weakly_dominant_rows():
    """(I / (gamma h) - J) for large steps: rows of a conservative diffusion operator plus a small diagonal.  The
    partitioned solve must not lose accuracy against the whole-system cyclic reduction."""
    rng = np.random.default_rng(7)
    n = 253
    for eps in (1e-2, 1e-6, 1e-10):
        for _ in range(10):
            kap = 10 ** rng.uniform(-2, 2, n + 1)
            a = -kap[:-1].copy(); c = -kap[1:].copy(); a[0] = 0; c[-1] = 0
            b = -(a + c) + eps * 10 ** rng.uniform(-2, 2, n) + kap[0] * (np.arange(n) == 0)
            d = rng.normal(size=n)
            xe = np.linalg.solve(_dense(a, b, c), d)
            ep = np.max(np.abs(partitioned_solve(n, a, b, c, d, 8) - xe)) / np.max(np.abs(xe))
            ew = np.max(np.abs(whole_system_cr(n, a, b, c, d) - xe)) / np.max(np.abs(xe))
            assert ep < 10.0 * ew + 1e-13, (eps, ep, ew)


def test_fixed_surface_row_and_decoupled_chunk():
    """fix_atmosphere_temperature_when_accreting zeroes row 0's couplings (and row 1's lower one); a chunk whose first row
    has no lower coupling has v = 0 and r_v = 0 / 0 -> 0."""
    rng = np.random.default_rng(3)
    n = 100
    a = -rng.uniform(0.1, 1, n); c = -rng.uniform(0.1, 1, n); a[0] = 0; c[-1] = 0
    b = np.abs(a) + np.abs(c) + 0.5
    c[0] = 0.0; a[1] = 0.0; b[0] = 4.0
    a[32] = 0.0    # chunk 1 decoupled from chunk 0 on its lower side
    d = rng.normal(size=n)
    xe = np.linalg.solve(_dense(a, b, c), d)
    assert np.max(np.abs(partitioned_solve(n, a, b, c, d, 4) - xe)) < 1e-13 * np.max(np.abs(xe)) * 100
