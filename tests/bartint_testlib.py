"""Shared helpers for the test-suite (importable as `bartint_testlib`; tests/ is put on sys.path by conftest)."""
import numpy as np


def kat_forest(k):
    """KAT json -> FlatForest (product container; also accepted by the oracle)."""
    from bartint_b200 import FlatForest
    f = k["forest"]
    return FlatForest(f["S"], f["m"], f["d"], np.array(f["tree_offset"]), np.array(f["var"]), np.array(f["cut"]),
                      np.array(f["left"]), np.array(f["right"]), np.array(f["mu"], dtype=float),
                      np.array(f["cut_offsets"]), np.array(f["cut_values"], dtype=float), f["y_min"], f["y_max"])


def has_gpu() -> bool:
    try:
        import ctypes as C
        from bartint_b200 import _lib
        n = C.c_int32(0)
        _lib.load().bartint_device_count(C.byref(n))
        return n.value > 0
    except Exception:
        return False
