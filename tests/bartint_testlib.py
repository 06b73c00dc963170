"""Shared helpers for the test-suite (importable as `bartint_testlib`; tests/ is put on sys.path by conftest)."""
import os

import numpy as np


def kernel_kind_params():
    """Kernel configurations every GPU parity test runs under.  Per-point outputs (moments, the full matrix, leaf
    ordinals) come from the register-gather kernel (default for d <= 16) or the shared-memory table kernel
    (BARTINT_KERNEL=table; the default for d > 16); evaluations that want per-draw sums only go through the bit-sliced
    kernel unless it is switched off ("+nobs": BARTINT_BITSLICE=0, sums from the per-point kernels).  The counting
    modes of the gather kernel (BARTINT_COUNT_SPLITS=1 / 2: per-draw sums of the one- and two-split trees from 1-D /
    2-D cumulative code histograms) only exist without the bit-sliced kernel."""
    return ["gather", "table", "gather+nobs", "table+nobs", "gather+count1", "gather+count2"]


def apply_kernel_kind(param, monkeypatch):
    kind, _, opt = param.partition("+")
    monkeypatch.setenv("BARTINT_KERNEL", kind)
    monkeypatch.setenv("BARTINT_COUNT_SPLITS", opt[-1] if opt.startswith("count") else "0")
    monkeypatch.setenv("BARTINT_BITSLICE", "0" if opt else "1")
    return kind


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
