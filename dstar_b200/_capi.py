"""ctypes binding of the product C ABI (include/dstar_b200.h, libdstar_b200.so).

Loading fails loudly when the CUDA library is missing or no device is present: there is no CPU
fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSTAR_B200_LIB: another build of the same library (kernel experiments: tools/micro_scan.py)
LIB_PATH = os.environ.get("DSTAR_B200_LIB") or os.path.join(_HERE, "libdstar_b200.so")
DATA_DIR = os.path.join(_HERE, "data")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class DStarError(RuntimeError):
    pass


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def p(a, typ=dp):
    if a is None:
        return typ()
    return a.ctypes.data_as(typ)


class MicroCtrl(C.Structure):
    """dstar_micro_ctrl (subset of NScool/public/NScool_controls.inc)."""
    _fields_ = [
        ("eos_gamma_melt_pt", C.c_double), ("eos_rsi_melt_pt", C.c_double),
        ("eos_nuclide_abundance_threshold", C.c_double),
        ("use_neutron_conductivity", C.c_int32), ("use_superfluid_phonon_conductivity", C.c_int32),
        ("use_pcy_for_ee_scattering", C.c_int32), ("use_page_for_eQ_scattering", C.c_int32),
        ("rad_full_off_lgrho", C.c_double), ("rad_full_on_lgrho", C.c_double),
        ("use_crust_nu_pair", C.c_int32), ("use_crust_nu_photo", C.c_int32), ("use_crust_nu_plasma", C.c_int32),
        ("use_crust_nu_bremsstrahlung", C.c_int32), ("use_crust_nu_pbf", C.c_int32),
        ("turn_on_shell_Urca", C.c_int32),
        ("shell_Urca_luminosity_coeff", C.c_double), ("lgP_shell_Urca", C.c_double),
    ]

    @classmethod
    def defaults(cls):
        """NScool/defaults/controls.defaults:118-148."""
        return cls(-1.0, -1.0, -1.0, 0, 0, 0, 0, -1.0, -1.0, 1, 1, 1, 1, 1, 0, 1.0, 28.5)


class EvolveCtrl(C.Structure):
    """dstar_evolve_ctrl."""
    _fields_ = [
        ("fix_core_temperature", C.c_int32), ("make_inner_boundary_insulating", C.c_int32),
        ("fix_atmosphere_temperature_when_accreting", C.c_int32), ("turn_on_extra_heating", C.c_int32),
        ("suppress_first_step_output", C.c_int32), ("starting_number_for_profile", C.c_int32),
        ("maximum_number_of_models", C.c_int32),
        ("integration_tolerance", C.c_double), ("min_lg_temperature_integration", C.c_double),
        ("max_lg_temperature_integration", C.c_double), ("maximum_timestep", C.c_double),
    ]

    @classmethod
    def defaults(cls):
        """NScool/defaults/controls.defaults:6-30."""
        return cls(1, 0, 1, 0, 0, 0, 10000, 1.0e-4, 7.0, 9.5, 0.0)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DStarError(
                "CUDA extension %s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "dstar_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.dstar_last_error.restype = C.c_char_p
        L.dstar_batch_last_kernel_ms.restype = C.c_double
        L.dstar_batch_last_kernel_ms.argtypes = [C.c_void_p, C.c_int]
        L.dstar_batch_last_launch_count.argtypes = [C.c_void_p]
        L.dstar_ctx_destroy.argtypes = [C.c_void_p]
        L.dstar_batch_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc, what):
    if rc < 0 and rc <= -100:
        raise DStarError("%s failed (%d): %s" % (what, rc, lib().dstar_last_error().decode()))
    return rc


def exported_symbols():
    """Symbols the header declares (used by the CPU-side ABI test)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "dstar_b200.h")).read()
    names = set(re.findall(r"\b(dstar_[a-z0-9_]+)\s*\(", hdr))
    more = os.path.join(os.path.dirname(_HERE), "include", "dstar_nscool.h")
    if os.path.exists(more):
        names |= set(re.findall(r"\b(dstar_[A-Za-z0-9_]+)\s*\(", open(more).read()))
    return sorted(names)
