"""Host logic of the abuntime composition path (no GPU): the C++ preprocessing behind `dstar_abuntime_table` against
an independent numpy restatement of dStar_crust/preprocessor/src/abuntime.f (read :18-133, reduce :135-179,
extend :181-268) on a synthetic file of the reference's format."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

GRAVITY, MDOT = 1.85052e14, 8.8e4 * 0.3
HZ90 = ["n", "mg40", "mg44", "mg48", "si46", "si50", "si54", "s52", "s56", "s60", "ar56", "ar62", "ar66", "ca56", "ca68",
        "ti56", "ti88", "cr56", "fe56"]                                   # HZ90_comp.f:9-12
LAYERS = ["fe56", "cr56", "ti56", "ca56", "ar56", "s52", "si46", "mg40", "ca68", "ar62", "s56", "si50", "mg44", "ar66",
          "s60", "si54", "mg48", "ti88"]                                  # HZ90_comp.f:29-44
PT = [7.235e26, 9.569e27, 1.152e29, 4.747e29, 1.361e30, 1.980e30, 2.253e30, 2.637e30, 2.771e30, 3.216e30, 3.825e30,
      4.699e30, 6.043e30, 7.233e30, 9.238e30, 1.228e31, 1.602e31]
XN = [0, 0, 0, 0, 0, 4 / 56, 10 / 56, 16 / 56, 22 / 56, 50 / 112, 56 / 112, 62 / 112, 68 / 112, 158 / 224, 164 / 224,
      170 / 224, 176 / 224, 360 / 448]


def hz90_Y(lgP):
    A = {n: float("".join(c for c in n if c.isdigit()) or 1) for n in HZ90}
    Y = np.zeros((len(lgP), len(HZ90)))
    lgPt = np.log10(PT)
    for j, x in enumerate(lgP):
        layer = next((i for i in range(17) if x <= lgPt[i]), 17)
        Y[j, 0] = XN[layer]
        Y[j, HZ90.index(LAYERS[layer])] = (1.0 - XN[layer]) / A[LAYERS[layer]]
    return Y


def parse(path, dlgP):
    """read_abuntime: fixed-width records; a point is kept only if lgP advanced by more than dlgP."""
    with open(path) as f:
        lines = f.read().split("\n")
    nion = int(lines[0][1:6])
    per = 1 + (nion + 4) // 5 + 1
    lgP, Y, isos, last = [], [], None, -100.0
    i = 1
    while i + per - 1 < len(lines) and lines[i].strip():
        time = float(lines[i][1:19])
        names, vals = [], []
        for ln in lines[i + 1:i + per - 1]:
            for c in range(5):
                if len(names) < nion:
                    names.append(ln[1 + 15 * c:6 + 15 * c].strip()); vals.append(float(ln[6 + 15 * c:16 + 15 * c]))
        x = np.log10(GRAVITY * MDOT * time)
        if x > last + dlgP:
            last = x; lgP.append(x); Y.append(vals)
        isos = names
        i += per
    return isos, np.array(lgP), np.array(Y)


def reduce_extend(isos, lgP, Y, dlgP, thresh):
    charged = np.array([n != "n" for n in isos])
    above = Y > thresh * Y[:, charged].max(axis=1, keepdims=True)
    keep = above.any(axis=0)
    net = [n for n, k in zip(isos, keep) if k]
    Yr = Y[:, keep]
    full = list(net)
    for n in HZ90:
        if n not in full:
            full.append(n)
    lo, hi = lgP.min(), lgP.max()
    nlow = max(int((lo - 22.0) / dlgP), 1); nhigh = max(int((33.5 - hi) / dlgP), 1)
    out = np.concatenate([lo - dlgP * np.arange(nlow, 0, -1), lgP, hi + dlgP * np.arange(1, nhigh + 1)])
    Yab = np.concatenate([np.repeat(Yr[:1], nlow, 0), Yr, np.repeat(Yr[-1:], nhigh, 0)])
    beta = np.where(out <= hi, 0.0, np.where(out > hi + 0.2, 1.0, (out - hi) / 0.2))
    Yout = np.zeros((len(out), len(full)))
    Yout[:, :len(net)] = Yab * (1.0 - beta)[:, None]
    Yh = hz90_Y(out)
    for i, n in enumerate(HZ90):
        Yout[:, full.index(n)] += Yh[:, i] * beta
    return full, out, Yout


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    import make_abuntime
    path = str(tmp_path_factory.mktemp("ab") / "synth.data")
    make_abuntime.write(path, 601)
    return path


@pytest.mark.parametrize("dlgP,thresh", [(0.005, 0.01), (0.02, 0.2)])
def test_abuntime_preprocessing_matches_restatement(synth, dlgP, thresh):
    import dstar_b200 as ds
    names, Z, A, lgP, Y = ds.abuntime_table(synth, dlgP, thresh)
    isos, lgP0, Y0 = parse(synth, dlgP)
    full, lgP_ref, Y_ref = reduce_extend(isos, lgP0, Y0, dlgP, thresh)
    assert names == full
    assert len(isos) == 64 and (len(full) < 64 + 18 if thresh > 0.1 else True)
    np.testing.assert_allclose(lgP, lgP_ref, rtol=0, atol=1e-12)
    np.testing.assert_allclose(Y, Y_ref, rtol=1e-14, atol=0)
    assert np.all(np.diff(lgP) > 0) and lgP[0] < 22.0 + 1.01 * dlgP and lgP[-1] > 33.5 - 1.01 * dlgP
    # network bookkeeping: Z, A from the names; neutrons first as in the file
    assert names[0] == "n" and Z[0] == 0 and A[0] == 1 and Z[names.index("fe56")] == 26 and A[names.index("ti88")] == 88
    # deep in the extension the composition is pure HZ90 (last layer: ti88 + neutrons)
    last = Y[-1]
    assert last[names.index("ti88")] == pytest.approx((1 - 360 / 448) / 88) and last[names.index("n")] == pytest.approx(360 / 448)
    assert np.count_nonzero(last) == 2
    # the reduction drops trace nuclides without renormalising (the lookup renormalises, dStar_crust_lib.f:137-139):
    # with the default 1 % threshold the table still carries > 99 % of the mass everywhere
    if thresh <= 0.01:
        assert np.all(np.abs((Y * A[None, :]).sum(axis=1) - 1.0) < 1e-2)


def test_abuntime_errors(tmp_path):
    import dstar_b200 as ds
    with pytest.raises(ds.DStarError):
        ds.abuntime_table(str(tmp_path / "missing.data"))
    bad = tmp_path / "bad.data"
    bad.write_text("    3\n garbage\n")
    with pytest.raises(ds.DStarError):
        ds.abuntime_table(str(bad))
