"""Shared helpers for the test-suite (importable as `bartint_testlib`; tests/ is put on sys.path by conftest)."""
import os

import numpy as np


def kernel_kind_params():
    """Kernel configurations every GPU parity test runs under: the register-gather kernel (default for d <= 16), the
    shared-memory table kernel (BARTINT_KERNEL=table; the default for d > 16) and the gather kernel in counting mode
    (BARTINT_COUNT_SPLITS=1: per-draw sums of the single-split trees from 1-D code histograms).  Level 2 of the
    counting mode (two-split trees from 2-D histograms) was merged after the round's GPU budget was spent: its
    planner is covered on CPU (test_planner_cpu.py), its kernels join the GPU matrix with BARTINT_TEST_COUNT2=1."""
    kinds = ["gather", "table", "gather+count1"]
    if os.environ.get("BARTINT_TEST_COUNT2") == "1":
        kinds.append("gather+count2")
    return kinds


def apply_kernel_kind(param, monkeypatch):
    kind, _, count = param.partition("+")
    monkeypatch.setenv("BARTINT_KERNEL", kind)
    monkeypatch.setenv("BARTINT_COUNT_SPLITS", count[-1] if count else "0")
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
