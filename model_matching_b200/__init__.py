"""model_matching_b200 -- B200-native StoCS hot path (drop-in for kuwt/model_matching's
stocs::stocs_estimator online methods).  The product is libstocs_b200.so (CUDA, sm_100a) behind
the C ABI in include/stocs_b200.h; this package holds its sources (csrc/), the C++ host shim
and CLIs (host/), a ctypes binding (_capi) and the synthetic-workload generators (synth)."""
from ._capi import Context, StocsError, lib, LIB_PATH, SYMBOLS  # noqa: F401

__all__ = ["Context", "StocsError", "lib", "LIB_PATH", "SYMBOLS"]
