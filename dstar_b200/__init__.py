"""dstar_b200 -- B200-native (sm_100a, FP64) hot path of nworbde/dStar's NScool crust-cooling code.

The package holds only what the path needs: csrc/ (CUDA kernels, C ABI, host-side mirror of
NScool_lib), data/ (rebuilt read-only tables) and thin ctypes wrappers.  There is no CPU fallback:
importing works anywhere, but every computation needs libdstar_b200.so and a CUDA device.
"""
from ._capi import DATA_DIR, LIB_PATH, DStarError, EvolveCtrl, MicroCtrl  # noqa: F401
from .api import (Batch, Context, NScool, NScool_init, NScool_shutdown, create_models, evolve_models,  # noqa: F401
                  abuntime_table, host_interp_pm, last_kernel_ms, run_batch, table_sets)

__all__ = ["Batch", "Context", "NScool", "NScool_init", "NScool_shutdown", "create_models", "evolve_models",
           "host_interp_pm", "abuntime_table", "MicroCtrl", "EvolveCtrl", "DStarError", "DATA_DIR", "LIB_PATH", "last_kernel_ms", "table_sets", "run_batch"]
