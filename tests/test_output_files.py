"""LOGS output in the reference's formats (NScool_history.f:25-89, NScool_profile.f:27-106, NScool_terminal.f:17-57):
the files under tests/golden/output_sample/ were written by the product on a B200 (tools/make_output_fixture.py)
together with the arrays they were written from.  Here they are parsed with the reference's own reader,
/root/reference/tools/reader.py (dStarReader: header names on line 4, column names on line 9), when the reference
tree is present, and always with a parser of the same layout rules; the GPU test regenerates the files."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SAMPLE = os.path.join(HERE, "golden", "output_sample")
REF_READER = "/root/reference/tools/reader.py"


class PlainReader:
    """tools/reader.py's layout rules restated: meta names on line 4, values on line 5; column names on line 9."""

    def __init__(self, path):
        lines = open(path).read().split("\n")
        self.headers = lines[3].split()
        self.header = dict(zip(self.headers, [float(v) for v in lines[4].split()]))
        self.columns = lines[8].split()
        rows = [l.split() for l in lines[9:] if l.strip()]
        self.data = {c: np.array([float(r[i]) for r in rows]) for i, c in enumerate(self.columns)}


def _readers(path):
    out = [("plain", PlainReader(path))]
    if os.path.exists(REF_READER):
        spec = importlib.util.spec_from_file_location("dstar_reference_reader", REF_READER)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out.append(("reference", mod.dStarReader(path)))
    return out


def _col(kind, r, name):
    if kind == "plain":
        return r.data[name]
    return np.asarray(r.data[name], dtype=float)


def _hdr(kind, r, name):
    return r.header[name] if kind == "plain" else getattr(r.header, name)


def check_sample(directory, arrays):
    a = arrays
    for kind, r in _readers(os.path.join(directory, "history.data")):
        assert list(r.headers) == ["gravity", "total_mass", "total_radius", "core_mass", "core_radius", "core_temp"]
        assert set(r.columns) == {"model", "time", "lg_Mdot", "lg_Teff", "lg_Lsurf", "lg_Lnu", "lg_Lnuc"}
        assert abs(_hdr(kind, r, "gravity") / a["grav"] - 1.0) < 1e-6 and abs(_hdr(kind, r, "total_mass") - a["Mtotal"]) < 1e-6
        assert abs(_hdr(kind, r, "core_temp") / a["Tcore"] - 1.0) < 1e-6
        np.testing.assert_array_equal(_col(kind, r, "model").astype(int), a["h_model"])
        t = _col(kind, r, "time")
        assert np.max(np.abs(t - a["h_time"]) / np.maximum(np.abs(a["h_time"]), 1e-30)) < 1e-6      # es15.6
        for c in ("lg_Mdot", "lg_Teff", "lg_Lsurf", "lg_Lnu", "lg_Lnuc"):
            assert np.max(np.abs(_col(kind, r, c) - a["h_" + c])) < 6e-7, (kind, c)               # f15.6
    for i in range(int(a["n_profiles"])):
        model = int(a["p%d_model" % i])
        for kind, r in _readers(os.path.join(directory, "profile%d" % model)):
            assert "accretion_rate" in r.headers and "temperature" in r.columns and len(r.columns) == 20
            assert int(_hdr(kind, r, "model")) == model
            assert abs(_hdr(kind, r, "time") - a["p%d_time" % i]) <= 1e-6 * abs(a["p%d_time" % i])
            for c in ("temperature", "luminosity", "pressure", "density", "cp", "eps_nu", "K"):
                got = _col(kind, r, c); want = a["p%d_%s" % (i, c)]
                assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-6, (kind, c)
            assert np.max(np.abs(_col(kind, r, "Gamma") - a["p%d_Gamma" % i])) < 6e-7 * np.max(a["p%d_Gamma" % i])
            np.testing.assert_array_equal(_col(kind, r, "zone").astype(int), np.arange(1, len(a["p%d_cp" % i]) + 1))
    # manifest (NScool_profile.f:49-55, :96): one line per profile
    man = open(os.path.join(directory, "profiles")).read().split("\n")
    assert "base profile filename = profile" in man[0]
    assert [int(l.split()[0]) for l in man[1:] if l.strip()] == [int(a["p%d_model" % i]) for i in range(int(a["n_profiles"]))]
    # terminal log: 11 columns of width 15; every second model; a title line before every fifth row of an epoch
    lines = open(os.path.join(directory, "terminal.txt")).read().split("\n")
    rows = [l for l in lines if l.strip() and not l.split()[0] == "model"]
    assert all(len(l) == 165 for l in rows)
    models = [int(l.split()[0]) for l in rows]
    assert models == [int(m) for m in a["t_model"] if m % 2 == 0]
    sel = np.isin(a["t_model"], models)
    got = np.array([[float(v) for v in l.split()[1:]] for l in rows])
    assert np.max(np.abs(got[:, 0] - a["t_tsec"][sel]) / np.maximum(a["t_tsec"][sel], 1e-30)) < 1e-6
    assert np.max(np.abs(got[:, 6] - a["t_lgTmax"][sel])) < 6e-7 and np.max(np.abs(got[:, 8] - a["t_lgTmin"][sel])) < 6e-7
    assert np.max(np.abs(got[:, 7] - a["t_lgP_Tmax"][sel])) < 6e-7 and np.max(np.abs(got[:, 9] - a["t_lgP_Tmin"][sel])) < 6e-7
    titles = [i for i, l in enumerate(lines) if l.split()[:1] == ["model"]]
    assert len(titles) >= 2 and lines[titles[0] - 1] == "" and lines[titles[0] + 1] == ""


def test_committed_output_sample_parses():
    check_sample(SAMPLE, dict(np.load(os.path.join(SAMPLE, "arrays.npz"))))


@pytest.mark.gpu
def test_writers_round_trip(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import make_output_fixture
    make_output_fixture.main(str(tmp_path))
    fresh = dict(np.load(os.path.join(str(tmp_path), "arrays.npz")))
    check_sample(str(tmp_path), fresh)
    # and the run is the committed one: same history, row for row
    old = dict(np.load(os.path.join(SAMPLE, "arrays.npz")))
    np.testing.assert_array_equal(fresh["h_model"], old["h_model"])
    assert np.max(np.abs(fresh["h_lg_Teff"] - old["h_lg_Teff"])) < 1e-9
