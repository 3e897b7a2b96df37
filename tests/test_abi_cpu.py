"""CPU-side checks of the product library: it loads, exports every symbol the headers declare, its
host utilities agree with the oracle, and every compute entry point FAILS LOUDLY without a GPU
(no CPU fallback).  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dstar_b200 as ds
from dstar_b200 import _capi

HAVE_GPU = os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = _capi.exported_symbols()
    assert len(names) >= 45, names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_headers_cite_reference_interfaces():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for hdr in ("dstar_b200.h", "dstar_nscool.h"):
        txt = open(os.path.join(root, "include", hdr)).read()
        cites = re.findall(r"[A-Za-z_/]+\.f:\d+", txt)
        assert len(cites) >= 6, (hdr, cites)


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_gpu():
    L = _capi.lib()
    h = C.c_void_p()
    rc = L.dstar_ctx_create(0, C.byref(h))
    assert rc == -101
    assert b"no CPU fallback" in L.dstar_last_error()
    with pytest.raises(ds.DStarError):
        ds.Context(0)
    with pytest.raises(ds.DStarError):
        ds.NScool_init()


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "dstar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt and "liboracle" not in txt and "dstar_oracle" not in txt, f


def test_host_interp_pm_matches_oracle(oracle):
    rng = np.random.default_rng(7)
    x = np.sort(rng.uniform(0, 10, 60))
    y = np.cumsum(rng.normal(size=60))
    f_host = ds.host_interp_pm(x, y)
    f_orac = oracle.interp_pm(x, y)
    np.testing.assert_allclose(f_host, f_orac, rtol=0, atol=0)


def test_namelist_and_epochs(tmp_path):
    """NScool_setup on the reference's own example inlist syntax (examples/basic_run/inlist)."""
    inl = tmp_path / "inlist"
    inl.write_text("""&controls
    ! comment
    write_interval_for_history = 1
    integration_tolerance = 1.0d-4
    target_resolution_lnP = 0.05
    core_temperature = 1.4e8
    atmosphere_temperature_when_accreting = 2.4d8
    number_epochs = 2
    basic_epoch_Mdots = 1.0e17, 0.0
    basic_epoch_boundaries = 0.0,50.0,1050.0
    turn_on_extra_heating = .TRUE.
    which_neutron_1S0_gap = 'sfb03'
    extra_real_controls = 4.0e13, 30.0
    scale_sf_critical_temperatures(:) = 1.0
    fix_Qimp = .TRUE.
    Qimp = 1.0
/
""")
    L = _capi.lib()
    sid = L.dstar_alloc_NScool()
    assert sid >= 1
    assert L.dstar_NScool_setup(sid, str(inl).encode()) == 0
    assert L.dstar_NScool_set_control(sid, b"Qimp", b"4.0") == 0
    assert L.dstar_NScool_set_control(sid, b"basic_epoch_Mdots", b"1.0e17,8*0.0") == 0
    assert L.dstar_NScool_set_control(sid, b"no_such_control", b"1") != 0
    bad = tmp_path / "bad"
    bad.write_text("&controls\n  not_a_control = 3\n/\n")
    assert L.dstar_NScool_setup(sid, str(bad).encode()) < 0
    assert L.dstar_NScool_setup(sid, b"/nonexistent/inlist") < 0
    assert L.dstar_dealloc_NScool(sid) == 0
    assert L.dstar_dealloc_NScool(sid) < 0
