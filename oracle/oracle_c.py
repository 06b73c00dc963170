"""ctypes wrapper of oracle/oracle_c.c (test infrastructure + CPU baseline; see the header there)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle_np import as_ns

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "oracle_c.c")):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(path)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(load().oracle_max_threads())


def _forest_args(f):
    return (C.c_int32(f.S), C.c_int32(f.m), C.c_int32(f.d), _p(f.tree_offset), _p(f.var), _p(f.cut), _p(f.left),
            _p(f.right), _p(f.mu), _p(f.cut_offsets), _p(f.cut_values))


def predict_reduce(forest, X, nthreads=1, want=("draw_mean", "point_mean", "point_var")):
    f = as_ns(forest)
    X = np.asfortranarray(np.asarray(X, np.float64))
    N = X.shape[0]
    out = {}
    if "draw_mean" in want:
        out["draw_mean"] = np.zeros(f.S)
    if "point_mean" in want:
        out["point_mean"] = np.zeros(N)
    if "point_var" in want:
        out["point_var"] = np.zeros(N)
    if "full_pred" in want:
        out["full_pred"] = np.zeros((f.S, N), order="F")
    load().oracle_predict_reduce(*_forest_args(f), C.c_double(f.y_min), C.c_double(f.y_max), C.c_int32(f.route), _p(X),
                                 C.c_int64(N),
                                 C.c_int32(nthreads), _p(out.get("draw_mean")), _p(out.get("point_mean")),
                                 _p(out.get("point_var")), _p(out.get("full_pred")))
    return out


def leaf_ordinals(forest, X):
    f = as_ns(forest)
    X = np.asfortranarray(np.asarray(X, np.float64))
    out = np.zeros((f.S * f.m, X.shape[0]), np.int32)
    load().oracle_leaf_ordinals(C.c_int32(f.S), C.c_int32(f.m), _p(f.tree_offset), _p(f.var), _p(f.cut), _p(f.left),
                                _p(f.right), _p(f.cut_offsets), _p(f.cut_values), C.c_int32(f.route), _p(X), C.c_int64(X.shape[0]),
                                _p(out))
    return out


def exact_integrals(forest, measure="uniform", n_draws=None, nthreads=1, lo=None, hi=None):
    f = as_ns(forest)
    n = f.S if n_draws is None else int(n_draws)
    out = np.zeros(n)
    code = {"uniform": 0, "gaussian": 1, "exponential": 2, "gaussian_correct": 3}[measure]
    lo_a = None if lo is None else np.ascontiguousarray(np.broadcast_to(np.asarray(lo, np.float64), (f.d,)))
    hi_a = None if hi is None else np.ascontiguousarray(np.broadcast_to(np.asarray(hi, np.float64), (f.d,)))
    load().oracle_exact_integrals(*_forest_args(f), C.c_int32(code), _p(lo_a), _p(hi_a), C.c_int32(n),
                                  C.c_int32(nthreads), _p(out))
    return out
