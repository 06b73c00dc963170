"""ctypes binding of libbartint_cuda.so (include/bartint.h).  No torch types cross this boundary.

The product path has NO fallback: if the shared library is missing this module raises, and every
compute entry point returns BARTINT_ERR_NODEVICE when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BARTINT_LIB: load another build of the same library instead (a sanitizer build, an older build to compare plans with)
LIB_PATH = os.environ.get("BARTINT_LIB") or os.path.join(HERE, "libbartint_cuda.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)
c_float_p = C.POINTER(C.c_float)
handle_p = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header
SIGNATURES = {
    "bartint_version": (C.c_int, []),
    "bartint_forest_bitslice_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "bartint_last_error": (C.c_char_p, []),
    "bartint_device_count": (C.c_int, [c_int32_p]),
    "bartint_set_device": (C.c_int, [C.c_int]),
    "bartint_init": (C.c_int, [C.c_int]),
    "bartint_group_size": (C.c_int, []),
    "bartint_group_exchange": (C.c_char_p, []),
    "bartint_shutdown": (None, []),
    "bartint_forest_from_dbarts098": (C.c_int, [C.POINTER(C.c_char_p), c_double_p, c_double_p, C.c_int64, C.c_int32,
                                                c_int32_p, c_double_p, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                                handle_p]),
    "bartint_forest_from_soa": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_int64_p, c_int32_p, c_int32_p, c_int32_p,
                                          c_int32_p, c_double_p, c_int32_p, c_double_p, C.c_double, C.c_double,
                                          handle_p]),
    "bartint_set_host_threads": (C.c_int, [C.c_int]),
    "bartint_design_step_dbarts098": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int32, c_int32_p,
                                                c_double_p, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                                C.c_void_p, c_double_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                                c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "bartint_forest_from_soa_routed": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_int64_p, c_int32_p, c_int32_p,
                                                 c_int32_p, c_int32_p, c_double_p, c_int32_p, c_double_p, C.c_double,
                                                 C.c_double, C.c_int32, handle_p]),
    "bartint_forest_from_value_splits": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_int64_p, c_int32_p, c_double_p, c_int32_p,
                                                   c_int32_p, C.c_int32, C.c_double, C.c_double, handle_p]),
    "bartint_forest_cutpoints": (C.c_int, [C.c_void_p, c_int32_p, c_double_p]),
    "bartint_eval_points_full_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bartint_sum_rows_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bartint_topk_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bartint_topk": (C.c_int, [c_double_p, C.c_int64, C.c_int32, c_int64_p, c_double_p]),
    "bartint_forest_route": (C.c_int, [C.c_void_p]),
    "bartint_forest_from_wbart": (C.c_int, [C.c_char_p, C.c_int64, C.c_int32, c_int32_p, c_double_p, C.c_double,
                                            handle_p]),
    "bartint_forest_image_size": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "bartint_forest_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "bartint_forest_unpack": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, handle_p]),
    "bartint_forest_info": (C.c_int, [C.c_void_p, c_int64_p]),
    "bartint_forest_export_soa": (C.c_int, [C.c_void_p, c_int64_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                            c_double_p]),
    "bartint_forest_destroy": (None, [C.c_void_p]),
    "bartint_stage_matrix": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_int, handle_p]),
    "bartint_finalize_integrals_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bartint_staged_flush": (C.c_int, [C.c_void_p]),
    "bartint_staged_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, handle_p]),
    "bartint_staged_free": (None, [C.c_void_p]),
    "bartint_points_from_host": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int32, handle_p]),
    "bartint_points_from_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p,
                                             handle_p]),
    "bartint_points_generate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                          handle_p]),
    "bartint_points_count": (C.c_int64, [C.c_void_p]),
    "bartint_points_codes": (C.c_int, [C.c_void_p, c_uint8_p]),
    "bartint_points_destroy": (None, [C.c_void_p]),
    "bartint_eval_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, c_double_p, C.c_int64,
                                      c_double_p, c_double_p, c_double_p, c_double_p]),
    "bartint_eval": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int32, C.c_uint32, c_double_p, C.c_int64,
                               c_double_p, c_double_p, c_double_p, c_double_p]),
    "bartint_eval_points_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "bartint_merge_moments_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, c_int32_p, C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bartint_argmax_ties_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "bartint_leaf_indices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_int32_p]),
    "bartint_exact_integrals": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p, C.c_int32, c_double_p]),
    "bartint_exact_integrals_device": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "bartint_finalize_draws_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "bartint_finalize_integrals": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, c_double_p, c_double_p, c_double_p]),
    "bartint_last_eval_profile": (C.c_int, [C.c_void_p, C.c_int32, c_float_p, c_int32_p, c_int32_p]),
    "bartint_kernel_launch_count": (C.c_int64, []),
}

BARTINT_SCALED_UNITS = 0x1
BARTINT_FORCE_WALK = 0x2
BARTINT_NO_BITSLICE = 0x4
BARTINT_DEVICE_GROUP = -1
MEASURES = {"uniform": 0, "gaussian": 1, "exponential": 2, "gaussian_correct": 3}
STATUS_NAMES = {0: "OK", 1: "ERR_ARG", 2: "ERR_PARSE", 3: "ERR_CUDA", 4: "ERR_UNSUPPORTED", 5: "ERR_NOMEM",
                6: "ERR_NODEVICE"}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  bartint_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
