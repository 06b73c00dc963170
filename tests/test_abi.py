"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/bartint.h declares,
and -- with no GPU -- every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bartint_b200 import _lib, synth, Forest, BartIntError
from bartint_testlib import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(name="bartint.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bartint_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bartint.h but not exported"
    # and the python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    # the host-only test hooks live in their own header (and build flag), outside the drop-in surface
    import hosthooks
    hooks = header_functions("bartint_test_hooks.h")
    assert hooks and all(h.startswith("bartint_debug_") for h in hooks) and not any("debug" in n for n in names)
    assert sorted(hosthooks.HOOK_SIGNATURES) == hooks and all(hasattr(lib, h) for h in hooks)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.bartint_version() >= 100
    assert isinstance(lib.bartint_last_error(), bytes)


def test_bad_arguments_are_status_codes_not_crashes():
    lib = _lib.load()
    h = C.c_void_p()
    off = np.zeros(1, np.int64)
    cut_off = np.zeros(2, np.int32)
    # NULL output handle
    rc = lib.bartint_forest_from_soa(0, 0, 1, off.ctypes.data_as(_lib.c_int64_p), None, None, None, None, None,
                                     cut_off.ctypes.data_as(_lib.c_int32_p), None, 0.0, 1.0, None)
    assert rc == 1 and b"NULL" in lib.bartint_last_error()
    # negative sizes
    rc = lib.bartint_forest_from_soa(-1, 1, 1, off.ctypes.data_as(_lib.c_int64_p), None, None, None, None, None,
                                     cut_off.ctypes.data_as(_lib.c_int32_p), None, 0.0, 1.0, C.byref(h))
    assert rc == 1
    # destroy(NULL) is a no-op
    lib.bartint_forest_destroy(None)
    lib.bartint_points_destroy(None)
    assert lib.bartint_points_count(None) == -1


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    rng = np.random.default_rng(0)
    ff = synth.prior_forest(1, 2, 3, rng.random((10, 2)))
    with pytest.raises(BartIntError) as e:
        Forest.from_soa(ff)
    assert e.value.status == 6 and "no CPU path" in str(e.value)
    n = C.c_int32(-1)
    assert _lib.load().bartint_device_count(C.byref(n)) == 0 and n.value == 0


def test_malformed_inputs_rejected_before_touching_the_device():
    rng = np.random.default_rng(0)
    ff = synth.prior_forest(1, 2, 3, rng.random((10, 2)))
    bad = synth.FlatForest(**{**ff.__dict__})
    bad.right = bad.right.copy()
    internal = np.nonzero(bad.var >= 0)[0]
    if len(internal):
        bad.right[internal[0]] = 10_000      # child index out of range
        with pytest.raises(BartIntError) as e:
            Forest.from_soa(bad)
        assert e.value.status == 1
    # descending cut points
    bad2 = synth.FlatForest(**{**ff.__dict__})
    bad2.cut_values = bad2.cut_values[::-1].copy()
    with pytest.raises(BartIntError) as e:
        Forest.from_soa(bad2)
    assert e.value.status == 1
    # malformed tree string
    with pytest.raises(BartIntError) as e:
        Forest.from_dbarts(np.array([["0 1 ."]], dtype=object), np.zeros((2, 1, 1)), rng.random((2, 1)),
                           [np.array([0.3, 0.6])], 0.0, 1.0)
    assert e.value.status == 2
    with pytest.raises(BartIntError) as e:     # variable index outside the cut-point table
        Forest.from_dbarts(np.array([["5 1 .."]], dtype=object), np.zeros((2, 1, 1)), rng.random((2, 1)),
                           [np.array([0.3, 0.6])], 0.0, 1.0)
    assert e.value.status == 2


def test_corrupted_soa_forests_are_rejected_or_harmless():
    """Seeded corruption sweep over the pre-flattened input (child indices, variables, cut indices, tree offsets): the
    library either rejects the forest with a status code or the change was immaterial (a leaf's unused fields, another
    valid value) -- it never crashes, hangs or plans from a malformed tree.  Host-only hook, no device needed."""
    import numpy as np
    import hosthooks
    from bartint_b200 import BartIntError, FlatForest, synth
    rng = np.random.default_rng(99)
    rejected = 0
    for it in range(400):
        d, m, S = int(rng.integers(1, 6)), int(rng.integers(1, 6)), int(rng.integers(1, 3))
        ff = synth.prior_forest(it, S, m, rng.random((20, d)), numcut=int(rng.choice([3, 20, 100])))
        a = {k: np.array(getattr(ff, k)) for k in ("var", "cut", "left", "right", "mu", "tree_offset")}
        nn = len(a["var"])
        j, kind = int(rng.integers(0, nn)), int(rng.integers(0, 6))
        was_leaf = a["var"][j] < 0
        if kind == 0:
            a["left"][j] = int(rng.integers(-5, nn + 5))
        elif kind == 1:
            a["right"][j] = int(rng.integers(-5, nn + 5))
        elif kind == 2:
            a["var"][j] = int(rng.integers(-3, d + 3))
        elif kind == 3:
            a["cut"][j] = int(rng.integers(-3, 300))
        elif kind == 4:
            a["tree_offset"][int(rng.integers(1, len(a["tree_offset"])))] += int(rng.choice([-3, -2, -1, 1, 2, 3]))
        else:
            a["right"][j] = a["left"][j]
        bad = FlatForest(S, m, d, a["tree_offset"], a["var"], a["cut"], a["left"], a["right"], a["mu"], ff.cut_offsets,
                         ff.cut_values, -0.5, 0.5)
        try:
            hosthooks.plan_soa(bad)
        except BartIntError as e:
            assert e.status in (1, 2, 4)
            rejected += 1
            continue
        # accepted: an internal node must still be a valid split
        internal = a["var"] >= 0
        assert (a["var"][internal] < d).all()
        ncut = np.diff(np.asarray(ff.cut_offsets))
        assert (a["cut"][internal] >= 0).all() and (a["cut"][internal] < ncut[a["var"][internal]]).all()
        if kind in (0, 1, 5) and not was_leaf:
            assert a["left"][j] != a["right"][j]
    assert rejected > 100
