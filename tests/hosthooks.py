"""Wrappers of the HOST-ONLY test hooks (include/bartint_test_hooks.h: bartint_debug_export_dbarts098,
bartint_debug_export_wbart, bartint_debug_plan_soa; compiled in with -DBARTINT_TEST_HOOKS).  They run the exporter / the group planner exactly as forest creation does but return the
result instead of uploading it, so they work on a machine without a GPU.  Used by tests/ and tools/ only: nothing
on the product path calls them."""
from __future__ import annotations

import ctypes as C

import numpy as np

from bartint_b200 import _lib
from bartint_b200._lib import c_double_p, c_int32_p, c_int64_p, c_uint8_p

# name -> (restype, argtypes) of the hooks; applied to the loaded library on first use (the product binding table
# bartint_b200/_lib.py covers include/bartint.h only)
HOOK_SIGNATURES = {
    "bartint_debug_export_wbart": (C.c_int, [C.c_char_p, C.c_int64, C.c_int32, c_int32_p, c_double_p, C.c_int64, C.c_int64,
                                             c_int64_p, c_int64_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                             c_double_p]),
    "bartint_debug_export_dbarts098": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int32, c_int32_p,
                                                 c_double_p, C.c_int32, C.c_int32, C.c_int64, c_int64_p, c_int64_p,
                                                 c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_double_p]),
    "bartint_debug_plan_soa": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_int64_p, c_int32_p, c_int32_p, c_int32_p,
                                         c_int32_p, c_double_p, c_int32_p, c_double_p, C.c_int64, c_int64_p,
                                         C.POINTER(C.c_uint32), c_int32_p, c_double_p, c_uint8_p, c_int32_p, c_int32_p,
                                         c_int32_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32)]),
    "bartint_debug_export_dbarts098_walk": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int32, c_int32_p, c_double_p,
                                                      C.c_int32, C.c_int32, C.c_int64, c_int64_p, C.POINTER(C.c_uint32), c_int32_p,
                                                      c_int32_p, c_double_p]),
    "bartint_debug_export_value_splits": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_int64_p, c_int32_p, c_double_p, c_int32_p,
                                                    c_int32_p, C.c_double, c_int64_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                                    c_double_p, c_int32_p, c_double_p]),
}


def _load():
    lib = _lib.load()
    for name, (res, args) in HOOK_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def export_dbarts(wire, tree_fits, x_train, cut_offsets, cut_values, want_arrays: bool = True):
    """Exporter + validation + walk-form packer + planner on the host.  wire: WireStrings.  Returns (n_nodes, dict of
    the flattened forest or None)."""
    lib = _load()
    x_train = np.asfortranarray(np.asarray(x_train, np.float64))
    n, d = x_train.shape
    m, S = wire.m, wire.S
    fits = np.asfortranarray(np.asarray(tree_fits, np.float64).reshape(n, m, S))
    coff = np.ascontiguousarray(cut_offsets, np.int32)
    cval = np.ascontiguousarray(cut_values, np.float64)
    nn = C.c_int64(0)
    cap, out = 1 << 62, None
    if want_arrays:
        cap = sum(max(2 * s.count(b".") - 1, 0) for s in wire._enc[:m * S]) if hasattr(wire, "_enc") else 0
        if getattr(wire, "_parent", None) is not None:         # a view: count from the C strings themselves
            cap = sum(max(2 * (wire.array[k] or b"").count(b".") - 1, 0) for k in range(m * S))
        out = dict(tree_offset=np.zeros(m * S + 1, np.int64), var=np.zeros(max(cap, 1), np.int32),
                   cut=np.zeros(max(cap, 1), np.int32), left=np.zeros(max(cap, 1), np.int32),
                   right=np.zeros(max(cap, 1), np.int32), mu=np.zeros(max(cap, 1), np.float64))
    o = out or {}
    rc = lib.bartint_debug_export_dbarts098(
        C.cast(wire.array, C.c_void_p), _p(fits, C.c_double), _p(x_train, C.c_double), n, d, _p(coff, C.c_int32),
        _p(cval, C.c_double), m, S, cap, C.byref(nn), _p(o.get("tree_offset"), C.c_int64), _p(o.get("var"), C.c_int32),
        _p(o.get("cut"), C.c_int32), _p(o.get("left"), C.c_int32), _p(o.get("right"), C.c_int32),
        _p(o.get("mu"), C.c_double))
    if rc != 0:
        from bartint_b200.forest import BartIntError
        raise BartIntError(rc, lib.bartint_last_error().decode())
    if out is not None:
        for k in ("var", "cut", "left", "right", "mu"):
            out[k] = out[k][:nn.value]
        out.update(S=S, m=m, d=d)
    return nn.value, out


def export_wbart(trees_text, cut_lists):
    """Host phases of Forest.from_wbart (BART::wbart `treedraws$trees` text).  Returns the flattened forest."""
    lib = _load()
    text = trees_text.encode() if isinstance(trees_text, str) else bytes(trees_text)
    d = len(cut_lists)
    coff = np.zeros(d + 1, np.int32)
    coff[1:] = np.cumsum([len(c) for c in cut_lists])
    cval = np.ascontiguousarray(np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0))
    cap_nodes = text.count(b"\n") + 2                    # one line per node at most
    cap_trees = cap_nodes
    dims = np.zeros(3, np.int64)
    out = dict(tree_offset=np.zeros(cap_trees + 1, np.int64), var=np.zeros(cap_nodes, np.int32),
               cut=np.zeros(cap_nodes, np.int32), left=np.zeros(cap_nodes, np.int32), right=np.zeros(cap_nodes, np.int32),
               mu=np.zeros(cap_nodes, np.float64))
    rc = lib.bartint_debug_export_wbart(text, len(text), d, _p(coff, C.c_int32), _p(cval, C.c_double), cap_trees,
                                        cap_nodes, _p(dims, C.c_int64), _p(out["tree_offset"], C.c_int64),
                                        _p(out["var"], C.c_int32), _p(out["cut"], C.c_int32), _p(out["left"], C.c_int32),
                                        _p(out["right"], C.c_int32), _p(out["mu"], C.c_double))
    if rc != 0:
        from bartint_b200.forest import BartIntError
        raise BartIntError(rc, lib.bartint_last_error().decode())
    S, m, nn = (int(v) for v in dims)
    out["tree_offset"] = out["tree_offset"][:S * m + 1]
    for k in ("var", "cut", "left", "right", "mu"):
        out[k] = out[k][:nn]
    out.update(S=S, m=m, d=d)
    return out


def plan_soa(ff) -> dict:
    """The planner's output for a FlatForest: what forest creation would upload for the fast kernels."""
    lib = _load()
    S, m, d = int(ff.S), int(ff.m), int(ff.d)
    n_trees = S * m
    toff = np.ascontiguousarray(ff.tree_offset, np.int64)
    n_nodes = int(toff[-1]) if n_trees else 0
    cap = n_trees + 2 * S + 1
    var, cut, left, right = (np.ascontiguousarray(getattr(ff, k), np.int32) for k in ("var", "cut", "left", "right"))
    mu = np.ascontiguousarray(ff.mu, np.float64)
    coff, cval = np.ascontiguousarray(ff.cut_offsets, np.int32), np.ascontiguousarray(ff.cut_values, np.float64)
    info = np.zeros(8, np.int64)
    out = dict(
        walk_nodes=np.zeros(2 * max(n_nodes, 1), np.uint32), leaf_off=np.zeros(n_trees + 1, np.int32),
        leaf_mu=np.zeros(max((n_nodes + n_trees) // 2, 1), np.float64), node_bit=np.zeros(max(n_nodes, 1), np.uint8),
        tree_perm=np.zeros(max(n_trees, 1), np.int32), draw_group_off=np.zeros(S + 1, np.int32),
        group_tree_off=np.zeros(cap + 1, np.int32), hdr=np.zeros(cap, np.uint32), split=np.zeros(cap, np.uint32),
        desc=np.zeros(16 * cap, np.uint32))
    rc = lib.bartint_debug_plan_soa(
        S, m, d, _p(toff, C.c_int64), _p(var, C.c_int32), _p(cut, C.c_int32), _p(left, C.c_int32), _p(right, C.c_int32),
        _p(mu, C.c_double), _p(coff, C.c_int32), _p(cval, C.c_double), cap, _p(info, C.c_int64),
        _p(out["walk_nodes"], C.c_uint32), _p(out["leaf_off"], C.c_int32), _p(out["leaf_mu"], C.c_double),
        _p(out["node_bit"], C.c_uint8), _p(out["tree_perm"], C.c_int32), _p(out["draw_group_off"], C.c_int32),
        _p(out["group_tree_off"], C.c_int32), _p(out["hdr"], C.c_uint32), _p(out["split"], C.c_uint32),
        _p(out["desc"], C.c_uint32))
    if rc != 0:
        from bartint_b200.forest import BartIntError
        raise BartIntError(rc, lib.bartint_last_error().decode())
    table_ok, gather, nb, regs, level, n_groups, dw, n_leaves = (int(v) for v in info)
    out.update(S=S, m=m, d=d, table_ok=bool(table_ok), gather=bool(gather), nb=nb, regs=regs, count_level=level,
               n_groups=n_groups, dw=dw, n_leaves=n_leaves, tree_off=toff.astype(np.int64), n_nodes=n_nodes)
    out["nodes_x"] = out["walk_nodes"][0::2][:n_nodes]
    out["nodes_y"] = out["walk_nodes"][1::2][:n_nodes]
    out["desc"] = out["desc"][:n_groups * dw].reshape(n_groups, dw) if n_groups else np.zeros((0, max(dw, 1)), np.uint32)
    for k in ("hdr", "split"):
        out[k] = out[k][:n_groups]
    out["group_tree_off"] = out["group_tree_off"][:n_groups + 1]
    return out


def export_value_splits(S, m, d, tree_offset, var, value, left=None, right=None, leaf_scale=1.0) -> dict:
    """The value-split exporter on the host: canonical pre-order SoA with cut indices + the cut lists it built."""
    lib = _load()
    toff = np.ascontiguousarray(tree_offset, np.int64)
    var = np.ascontiguousarray(var, np.int32)
    value = np.ascontiguousarray(value, np.float64)
    lft = None if left is None else np.ascontiguousarray(left, np.int32)
    rgt = None if right is None else np.ascontiguousarray(right, np.int32)
    nn = int(toff[-1])
    out = dict(tree_offset=np.zeros(S * m + 1, np.int64), var=np.zeros(max(nn, 1), np.int32), cut=np.zeros(max(nn, 1), np.int32),
               left=np.zeros(max(nn, 1), np.int32), right=np.zeros(max(nn, 1), np.int32), mu=np.zeros(max(nn, 1)),
               cut_offsets=np.zeros(d + 1, np.int32), cut_values=np.zeros(max(nn, 1)))
    rc = lib.bartint_debug_export_value_splits(S, m, d, _p(toff, C.c_int64), _p(var, C.c_int32), _p(value, C.c_double),
                                               _p(lft, C.c_int32), _p(rgt, C.c_int32), float(leaf_scale),
                                               _p(out["tree_offset"], C.c_int64), _p(out["var"], C.c_int32), _p(out["cut"], C.c_int32),
                                               _p(out["left"], C.c_int32), _p(out["right"], C.c_int32), _p(out["mu"], C.c_double),
                                               _p(out["cut_offsets"], C.c_int32), _p(out["cut_values"], C.c_double))
    if rc != 0:
        from bartint_b200.forest import BartIntError
        raise BartIntError(rc, lib.bartint_last_error().decode())
    for k in ("var", "cut", "left", "right", "mu"):
        out[k] = out[k][:nn]
    out["cut_values"] = out["cut_values"][:int(out["cut_offsets"][-1])]
    return out


def export_dbarts_walk(wire, tree_fits, x_train, cut_offsets, cut_values, node_capacity) -> dict:
    """The one-pass exporter of the fused design step on the host: walk form + statistics."""
    lib = _load()
    x_train = np.asfortranarray(np.asarray(x_train, np.float64))
    n, d = x_train.shape
    m, S = wire.m, wire.S
    fits = np.asfortranarray(np.asarray(tree_fits, np.float64).reshape(n, m, S))
    coff = np.ascontiguousarray(cut_offsets, np.int32)
    cval = np.ascontiguousarray(cut_values, np.float64)
    stats = np.zeros(8, np.int64)
    cap = max(int(node_capacity), 1)
    out = dict(walk_nodes=np.zeros(2 * cap, np.uint32), tree_off=np.zeros(m * S + 1, np.int32),
               leaf_off=np.zeros(m * S + 1, np.int32), leaf_mu=np.zeros(cap, np.float64))
    rc = lib.bartint_debug_export_dbarts098_walk(wire.array, _p(fits, C.c_double), _p(x_train, C.c_double), n, d,
                                                 _p(coff, C.c_int32), _p(cval, C.c_double), m, S, cap, _p(stats, C.c_int64),
                                                 _p(out["walk_nodes"], C.c_uint32), _p(out["tree_off"], C.c_int32),
                                                 _p(out["leaf_off"], C.c_int32), _p(out["leaf_mu"], C.c_double))
    if rc != 0:
        from bartint_b200.forest import BartIntError
        raise BartIntError(rc, lib.bartint_last_error().decode())
    out["stats"] = stats
    out["walk_nodes"] = out["walk_nodes"][:2 * int(stats[0])]
    out["leaf_mu"] = out["leaf_mu"][:int(stats[1])]
    return out
