"""Synthetic posteriors and Genz-shaped data (tests + benchmark inputs; not on the evaluation path).

The BART MCMC fit (`bart(...)`, BARTInt.R:256,282,363; bartMean.R:41-42) is out of scope and cannot run
here (no R).  Its *output* -- S draws x m trees over a cut-point grid built from the training design --
is synthesised from the BART tree prior instead, in the same
flattened form the exporter produces from a real dbarts state, and optionally in dbarts' own
0.9-8 wire format (tree strings + savedTreeFits) so the exporter is exercised on it.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from .forest import FlatForest

HERE = os.path.dirname(os.path.abspath(__file__))
_synth = None


def _load_synth():
    global _synth
    if _synth is None:
        from .build import build_synth
        lib = C.CDLL(build_synth())
        lib.bartint_synth_create.restype = C.c_void_p
        lib.bartint_synth_create.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                             C.c_double, C.c_double, C.c_double, C.c_int32]
        lib.bartint_synth_nodes.restype = C.c_int64
        lib.bartint_synth_nodes.argtypes = [C.c_void_p]
        lib.bartint_synth_copy.restype = None
        lib.bartint_synth_copy.argtypes = [C.c_void_p] * 7
        lib.bartint_synth_free.restype = None
        lib.bartint_synth_free.argtypes = [C.c_void_p]
        _synth = lib
    return _synth


# ---------------------------------------------------------------------------------------------
# cut points and codes
# ---------------------------------------------------------------------------------------------
def uniform_cutpoints(x_train: np.ndarray, numcut: int = 100):
    """dbarts cut points with usequants=FALSE:
    cut[j][k] = min_j + (k+1) * (max_j - min_j) / (numcut + 1), k = 0..numcut-1, over the training column."""
    x_train = np.asarray(x_train, np.float64)
    out = []
    for j in range(x_train.shape[1]):
        lo, hi = x_train[:, j].min(), x_train[:, j].max()
        out.append(lo + (np.arange(numcut) + 1.0) * (hi - lo) / (numcut + 1.0))
    return out


def bin_codes(X: np.ndarray, cut_lists) -> np.ndarray:
    """(N, d) integer cut map: #{k : x > cut[j][k]} (host helper for building synthetic forests)."""
    X = np.asarray(X, np.float64)
    out = np.zeros(X.shape, np.uint8)
    for j, cuts in enumerate(cut_lists):
        out[:, j] = np.where(np.isnan(X[:, j]), 0, np.searchsorted(cuts, X[:, j], side="left"))
    return out


def _pack_cuts(cut_lists):
    d = len(cut_lists)
    off = np.zeros(d + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in cut_lists])
    vals = np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0)
    return off, vals


# ---------------------------------------------------------------------------------------------
# forests
# ---------------------------------------------------------------------------------------------
def prior_forest(seed: int, S: int, m: int, x_train: np.ndarray, y_train=None, numcut: int = 100, k: float = 2.0,
                 base: float = 0.95, power: float = 2.0, max_depth: int = 32, cut_lists=None) -> FlatForest:
    """S x m trees from the BART tree prior on the cut map of x_train (every leaf owns a training row)."""
    x_train = np.asarray(x_train, np.float64)
    n, d = x_train.shape
    if cut_lists is None:
        cut_lists = uniform_cutpoints(x_train, numcut)
    codes = np.ascontiguousarray(bin_codes(x_train, cut_lists).T)   # feature-major [j, i]
    lib = _load_synth()
    sigma_mu = 0.5 / (k * math.sqrt(max(m, 1)))
    h = lib.bartint_synth_create(seed, S, m, d, n, codes.ctypes.data, base, power, sigma_mu, max_depth)
    if not h:
        raise MemoryError("synthetic forest generation failed")
    try:
        nn = lib.bartint_synth_nodes(h)
        off = np.zeros(S * m + 1, np.int64)
        var, cut, left, right = (np.zeros(nn, np.int32) for _ in range(4))
        mu = np.zeros(nn)
        lib.bartint_synth_copy(h, off.ctypes.data, var.ctypes.data, cut.ctypes.data, left.ctypes.data,
                               right.ctypes.data, mu.ctypes.data)
    finally:
        lib.bartint_synth_free(h)
    cut_offsets, cut_values = _pack_cuts(cut_lists)
    y_min, y_max = (-0.5, 0.5) if y_train is None else (float(np.min(y_train)), float(np.max(y_train)))
    return FlatForest(S, m, d, off, var, cut, left, right, mu, cut_offsets, cut_values, y_min, y_max,
                      meta={"seed": seed, "generator": "bart_prior"})


def forest_from_trees(trees, S: int, m: int, cut_lists, y_min=-0.5, y_max=0.5) -> FlatForest:
    """Build a FlatForest from nested python trees: leaf = float mu, internal = (var, cut_idx, left, right)."""
    assert len(trees) == S * m
    off = [0]
    V, Cc, L, R, MU = [], [], [], [], []
    for tree in trees:
        var, cut, left, right, mu = [], [], [], [], []

        def rec(node):
            me = len(var)
            if isinstance(node, tuple):
                v, c, l, r = node
                var.append(int(v)); cut.append(int(c)); left.append(-1); right.append(-1); mu.append(0.0)
                left[me] = rec(l)
                right[me] = rec(r)
            else:
                var.append(-1); cut.append(0); left.append(-1); right.append(-1); mu.append(float(node))
            return me

        import sys
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
        rec(tree)
        V += var; Cc += cut; L += left; R += right; MU += mu
        off.append(off[-1] + len(var))
    cut_offsets, cut_values = _pack_cuts(cut_lists)
    return FlatForest(S, m, len(cut_lists), np.array(off), np.array(V, np.int32), np.array(Cc, np.int32),
                      np.array(L, np.int32), np.array(R, np.int32), np.array(MU), cut_offsets, cut_values, y_min, y_max)


def adversarial_forest(kind: str, S: int, m: int, d: int, numcut: int = 100, seed: int = 0, depth: int = 12) -> FlatForest:
    """Hand-shaped trees the prior rarely produces: 'stumps', 'left_chain', 'right_chain', 'same_var',
    'balanced', 'edge_cuts' (cut indices 0 and numcut-1)."""
    rng = np.random.default_rng(seed)
    cut_lists = [np.linspace(0.0, 1.0, numcut + 2)[1:-1] for _ in range(d)]

    def mu():
        return float(rng.normal(0, 0.05))

    def chain(depth_left, right_side):
        node = mu()
        for _ in range(depth_left):
            v, c = int(rng.integers(d)), int(rng.integers(numcut))
            node = (v, c, mu(), node) if right_side else (v, c, node, mu())
        return node

    def same_var(depth_left, v, lo, hi):
        if depth_left == 0 or hi - lo < 1:
            return mu()
        c = int(rng.integers(lo, hi))
        return (v, c, same_var(depth_left - 1, v, lo, c), same_var(depth_left - 1, v, c + 1, hi))

    def balanced(depth_left):
        if depth_left == 0:
            return mu()
        return (int(rng.integers(d)), int(rng.integers(numcut)), balanced(depth_left - 1), balanced(depth_left - 1))

    trees = []
    for _ in range(S * m):
        if kind == "stumps":
            trees.append(mu())
        elif kind == "left_chain":
            trees.append(chain(depth, False))
        elif kind == "right_chain":
            trees.append(chain(depth, True))
        elif kind == "same_var":
            trees.append(same_var(min(depth, 5), int(rng.integers(d)), 0, numcut))
        elif kind == "balanced":
            trees.append(balanced(min(depth, 3)))
        elif kind == "edge_cuts":
            trees.append((int(rng.integers(d)), 0, mu(), (int(rng.integers(d)), numcut - 1, mu(), mu())))
        else:
            raise ValueError(kind)
    return forest_from_trees(trees, S, m, cut_lists)


# ---------------------------------------------------------------------------------------------
# integrands (restated formulas, used only to make "Genz-shaped" training responses)
# ---------------------------------------------------------------------------------------------
def genz_gaussian(X):
    """Gaussian peak family, a = 100/d^2, u = 0.5 (formula of src/genz/genz.R:192-242)."""
    X = np.atleast_2d(X)
    d = X.shape[1]
    a = 100.0 / d ** 2
    return np.exp(-((X - 0.5) ** 2 * a ** 2).sum(axis=1))


def genz_step(X, regularizer=1e-5):
    """Step function with one jump at x1 = 1/2 (final values of src/genz/genz.R:350-384 for jumps=1)."""
    X = np.atleast_2d(X)
    return np.where(X[:, 0] <= 0.5, regularizer, 1.0 - regularizer)


def portfolio_loss(X, gamma=2.0, regularizer=1e-6):
    """Rare-event indicator (formula of src/rareFunctions.R:47-83)."""
    X = np.atleast_2d(X)
    d = X.shape[1]
    a = 0.5 * np.arange(1, d + 1)
    c = 0.2 * np.arange(1, d + 1)
    return ((c * (X > a)).sum(axis=1) > gamma).astype(np.float64) + regularizer


def sample_measure(rng, measure: str, n: int, d: int) -> np.ndarray:
    """Points from the integration measure as the reference draws candidates (BARTInt.R:300-309)."""
    if measure == "uniform":
        return rng.random((n, d))
    if measure == "exponential":
        return rng.exponential(1.0, (n, d))
    if measure == "gaussian":   # N(0.5, 1) truncated to [0, 1] by inverse CDF
        from scipy.special import ndtr, ndtri
        lo, hi = ndtr(-0.5), ndtr(0.5)
        return 0.5 + ndtri(lo + (hi - lo) * rng.random((n, d)))
    raise ValueError(measure)


def survey_design(rng, n: int):
    """PUMS-shaped covariates: 6 categorical codes + ordinal Education with the reference's 1e-5 jitter
    (poptMean_trained_bin.R:66), padded to d = 10 numeric columns."""
    cards = (6, 2, 2, 2, 2, 9)
    cols = [rng.integers(1, c + 1, n).astype(np.float64) for c in cards]
    cols.append(rng.integers(1, 25, n).astype(np.float64) + 1e-5)
    while len(cols) < 10:
        cols.append(rng.integers(0, 2, n).astype(np.float64))
    X = np.stack(cols, axis=1)
    p = 1.0 / (1.0 + np.exp(-(0.25 * (X[:, 6] - 12.0) + 0.3 * (X[:, 0] - 3.0))))
    y = np.where(rng.random(n) < p, 1.0 - 1e-5, 1e-5)      # poptMean_trained_bin.R:74-77
    return X, y


# ---------------------------------------------------------------------------------------------
# BASELINE.json configurations
# ---------------------------------------------------------------------------------------------
CONFIGS = {
    1: dict(name="cfg1 integrationMain.R 1 2 1 matern32 1 uniform 1 1", d=1, n=20, m=50, S=2000, N=0, C=100,
            measure="uniform", integrand="genz_cont"),
    2: dict(name="cfg2 Genz Gaussian-peak d=10, 50 trees x 1000 draws, 1e6 uniform MC points", d=10, n=200, m=50,
            S=1000, N=1_000_000, C=100, measure="uniform", integrand="genz_gaussian"),
    3: dict(name="cfg3 Genz step d=20, 200 trees x 2000 draws, 1e7 MC points + exact integral", d=20, n=400, m=200,
            S=2000, N=10_000_000, C=100, measure="uniform", integrand="genz_step"),
    4: dict(name="cfg4 rare-event portfolio_loss d=10, 1e8 Gaussian-measure points", d=10, n=5000, m=50, S=2000,
            N=100_000_000, C=100, measure="gaussian", integrand="portfolio_loss"),
    5: dict(name="cfg5 survey bartMean, 2e5 candidates x 10 features, 1000 draws", d=10, n=20, m=50, S=1000,
            N=0, C=200_000, measure="survey", integrand="survey"),
}


def make_design(cfg: int, seed: int | None = None, scale: float = 1.0):
    """Training design + response for a BASELINE config (scale shrinks n for small tests)."""
    c = CONFIGS[cfg]
    rng = np.random.default_rng(np.random.PCG64(cfg if seed is None else seed))
    n = max(2, int(round(c["n"] * scale)))
    if c["integrand"] == "survey":
        return survey_design(rng, n)
    X = sample_measure(rng, c["measure"], n, c["d"])
    if c["integrand"] == "genz_gaussian":
        y = genz_gaussian(X)
    elif c["integrand"] == "genz_step":
        y = genz_step(X)
    elif c["integrand"] == "portfolio_loss":
        y = portfolio_loss(X)
    else:   # Genz "continuous" family, a = 150/d^3, u = 0.5
        a = 150.0 / c["d"] ** 3
        y = np.exp(-(a * np.abs(X - 0.5)).sum(axis=1))
    return X, y


def make_points(cfg: int, n_points: int, seed: int = 12345):
    """Evaluation points for a config: MC points from the measure, or survey candidates."""
    c = CONFIGS[cfg]
    rng = np.random.default_rng(np.random.PCG64(seed))
    if c["integrand"] == "survey":
        return survey_design(rng, n_points)[0]
    return sample_measure(rng, c["measure"], n_points, c["d"])
