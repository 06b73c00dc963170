"""CPU ORACLE for the BART-Int posterior-evaluation hot path (numpy restatement).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``bart-int_b200/`` imports or calls it.

PARITY UNPINNED.  The reference (/root/reference, ImperialCollegeLondon/BART-Int) is pure R
and ships no tests, golden vectors, saved posteriors or prediction dumps for this path
(SURVEY.md section 8c); R, dbarts 0.9-8, data.tree, matrixStats and mvtnorm are not
installable here.  The arithmetic that lives in-tree (src/BARTInt.R:10-223, 260-262, 312-315;
src/survey_design/bartMean.R:47-51, 74-82) is restated line by line below.  The arithmetic that
lives in the un-vendored third-party packages is restated from their published behaviour and is
tagged:

  [dbarts-mem]      dbarts 0.9-8 (README.md:83,102-106 pins the version): tree-string grammar,
                    routing rule ``x <= cut -> left``, cut-point formula, response scaling
                    ``y' = (y - ymin)/(ymax - ymin) - 0.5``, predict() returning S x N.
  [matrixStats-mem] matrixStats::colVars: unbiased (n-1) column variance.
  [mvtnorm-mem]     mvtnorm::pmvnorm in one dimension reduces to pnorm(upper) - pnorm(lower) of
                    the standardised bounds ("univariate: using pnorm").
  [R-mem]           base R: sum()/mean()/rowMeans() accumulate in long double; mean() adds one
                    refinement pass; pexp(q) = -expm1(-q) for q > 0 else 0.

The substitute pins (hand-computed known-answer tests, invariants, grid quadrature) live in
tests/golden/ and tests/test_oracle.py.  tools/dump_dbarts_fixture.R produces real dbarts
fixtures on a box that has R; until those exist every parity claim is against THIS restatement.

Forest container used throughout ("flat forest"): any object/dict exposing
    S, m, d                      draws, trees per draw, features
    tree_offset  int64[S*m+1]    first node of tree (s, t) at index t + s*m  (R: savedTrees[t, s])
    var          int32[nodes]    0-based split variable, -1 for a leaf
    cut          int32[nodes]    0-based cut-point index into that variable's cut list
    left, right  int32[nodes]    child node index RELATIVE to the tree's first node, -1 for a leaf
    mu           float64[nodes]  leaf value in dbarts' scaled units (0 for internal nodes)
    cut_offsets  int32[d+1], cut_values float64[cut_offsets[d]]   ascending per variable
    y_min, y_max float           range of the training response
    route        int (optional)  0 = dbarts rule "left iff x <= cut" (default); 1 = BART::wbart rule "left iff x < cut"
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

LD = np.longdouble  # 80-bit on x86-64: what R's LDOUBLE is on the reference's platforms
ROUTE_DBARTS = 0    # left iff x <= cut; NaN compares false -> left  [dbarts-mem]
ROUTE_BART = 1      # left iff x <  cut; NaN compares false -> right [BART-mem: tree::bn, "if(x[v] < xi[v][c]) left"]


def _get(forest, name):
    return forest[name] if isinstance(forest, dict) else getattr(forest, name)


def as_ns(forest) -> SimpleNamespace:
    """Normalise a flat forest (dict or object) into a namespace of contiguous arrays."""
    ns = SimpleNamespace()
    ns.S, ns.m, ns.d = int(_get(forest, "S")), int(_get(forest, "m")), int(_get(forest, "d"))
    ns.tree_offset = np.ascontiguousarray(_get(forest, "tree_offset"), dtype=np.int64)
    for k in ("var", "cut", "left", "right", "cut_offsets"):
        setattr(ns, k, np.ascontiguousarray(_get(forest, k), dtype=np.int32))
    ns.mu = np.ascontiguousarray(_get(forest, "mu"), dtype=np.float64)
    ns.cut_values = np.ascontiguousarray(_get(forest, "cut_values"), dtype=np.float64)
    ns.y_min, ns.y_max = float(_get(forest, "y_min")), float(_get(forest, "y_max"))
    try:
        ns.route = int(_get(forest, "route"))
    except (KeyError, AttributeError):
        ns.route = ROUTE_DBARTS
    return ns


# --------------------------------------------------------------------------------------
# a6 / a11  getTree: dbarts 0.9-8 saved-tree strings + savedTreeFits -> flat forest
# --------------------------------------------------------------------------------------

def tokenize_tree_string(s: str) -> list[str]:
    """BARTInt.R:118-119  strsplit(gsub("\\\\.", "\\\\. ", treeString), " ", fixed=TRUE)[[1]].

    R's strsplit drops a trailing empty token but keeps interior empty ones; dbarts strings
    never contain double spaces, so filtering empties is equivalent for well-formed input.
    """
    return [tok for tok in s.replace(".", ". ").split(" ") if tok != ""]


def parse_tree_tokens(tokens: list[str]):
    """dbarts:::buildTree [dbarts-mem]: pre-order grammar  tree := "." | VAR CUT tree tree.

    VAR and CUT are 0-based in the string (buildTree adds 1; BARTInt.R:45 then indexes
    cutPoints[[splitVar]][splitIndex] 1-based).  Returns pre-order arrays with child indices
    relative to the root (left child is always self+1 in pre-order).
    """
    var, cut, left, right = [], [], [], []
    pos = 0

    # iterative pre-order build (R recurses; result is identical)
    def build():
        nonlocal pos
        if pos >= len(tokens):
            raise ValueError("truncated tree string")
        tok = tokens[pos]
        me = len(var)
        if tok == ".":
            pos += 1
            var.append(-1); cut.append(0); left.append(-1); right.append(-1)
            return me
        v = int(tok); c = int(tokens[pos + 1]); pos += 2
        var.append(v); cut.append(c); left.append(-1); right.append(-1)
        left[me] = build()
        right[me] = build()
        return me

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 4 * len(tokens) + 100))
    try:
        build()
    finally:
        sys.setrecursionlimit(old)
    if pos != len(tokens):
        raise ValueError("trailing tokens after tree (tree$remainder non-empty)")
    return (np.array(var, np.int32), np.array(cut, np.int32),
            np.array(left, np.int32), np.array(right, np.int32))


def route_rows(var, cut, left, right, X, cut_lists):
    """dbarts:::fillObservationsForNode [dbarts-mem] + routing rule A4: left iff x <= cut value.

    X is (n, d) float64; returns the leaf node index (relative) of every row.
    """
    n = X.shape[0]
    node = np.zeros(n, np.int64)
    rows = np.arange(n)
    while True:
        v = var[node]
        internal = v >= 0
        if not internal.any():
            return node
        vi = np.where(internal, v, 0)
        cv = np.array([cut_lists[a][b] for a, b in zip(vi, np.where(internal, cut[node], 0))],
                      dtype=np.float64) if n else np.zeros(0)
        go_right = X[rows, vi] > cv          # NaN compares false -> left
        nxt = np.where(go_right, right[node], left[node])
        node = np.where(internal, nxt, node)


def forest_from_dbarts(tree_strings, tree_fits, x_train, cut_lists, y_min, y_max):
    """getTree for every (tree, draw): BARTInt.R:92-133.

    tree_strings : array-like of str, shape (m, S)        == state[[1]]@savedTrees[treeNum, sampleNum]  (:110)
    tree_fits    : float64 (n, m, S)                       == state[[1]]@savedTreeFits[, treeNum, sampleNum] (:111)
    x_train      : float64 (n, d)                          == sampler$data@x
    cut_lists    : list of d ascending float64 arrays      == dbarts:::createCutPoints(sampler) (:108)
    Leaf mu = treeFits[first training row routed to that leaf] (fillPlotInfoForNode [dbarts-mem], used :172).
    A leaf that receives no training row gets mu = NaN (dbarts never produces one).
    """
    tree_strings = np.asarray(tree_strings, dtype=object)
    m, S = tree_strings.shape
    x_train = np.asarray(x_train, np.float64)
    n, d = x_train.shape
    tree_fits = np.asarray(tree_fits, np.float64).reshape(n, m, S)
    offs = [0]
    V, C, L, R, MU = [], [], [], [], []
    for s in range(S):
        for t in range(m):
            var, cut, left, right = parse_tree_tokens(tokenize_tree_string(str(tree_strings[t, s])))
            leaf_of_row = route_rows(var, cut, left, right, x_train, cut_lists)
            mu = np.zeros(len(var))
            for node in np.nonzero(var < 0)[0]:
                rows = np.nonzero(leaf_of_row == node)[0]
                mu[node] = tree_fits[rows[0], t, s] if len(rows) else np.nan
            V.append(var); C.append(cut); L.append(left); R.append(right); MU.append(mu)
            offs.append(offs[-1] + len(var))
    cut_offsets = np.zeros(d + 1, np.int32)
    cut_offsets[1:] = np.cumsum([len(c) for c in cut_lists])
    return dict(S=S, m=m, d=d, tree_offset=np.array(offs, np.int64),
                var=np.concatenate(V), cut=np.concatenate(C), left=np.concatenate(L),
                right=np.concatenate(R), mu=np.concatenate(MU), cut_offsets=cut_offsets,
                cut_values=np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0),
                y_min=float(y_min), y_max=float(y_max))


# --------------------------------------------------------------------------------------
# a1  predict(model, X): dbarts C++ predict [dbarts-mem], call sites BARTInt.R:312,380; bartMean.R:47,74
# --------------------------------------------------------------------------------------

def bin_points(X, cut_offsets, cut_values, route=ROUTE_DBARTS):
    """dbarts' integer cut map of a test matrix [dbarts-mem]:  b = #{k : x > cut[j][k]}.

    Right iff b > splitIndex0  <=>  x > cut[j][splitIndex0] for ascending cuts.  NaN -> 0.
    route = ROUTE_BART: b = #{k : x >= cut[j][k]} (right iff x >= cut), NaN -> number of cuts (always right).
    """
    X = np.asarray(X, np.float64)
    N, d = X.shape
    codes = np.zeros((N, d), np.int32)
    for j in range(d):
        cuts = cut_values[cut_offsets[j]:cut_offsets[j + 1]]
        col = X[:, j]
        if route == ROUTE_BART:
            b = np.searchsorted(cuts, col, side="right")  # number of cuts <= x
            codes[:, j] = np.where(np.isnan(col), len(cuts), b)
        else:
            b = np.searchsorted(cuts, col, side="left")  # number of cuts strictly below x
            codes[:, j] = np.where(np.isnan(col), 0, b)
    return codes


def leaf_nodes(forest, X):
    """Leaf node index (relative to its tree) reached by every point in every tree.

    Returns int32 array (S*m, N), row index t + s*m.  Routing rule A4 on fp64 values.
    """
    f = as_ns(forest)
    X = np.asarray(X, np.float64)
    N = X.shape[0]
    out = np.zeros((f.S * f.m, N), np.int32)
    rows = np.arange(N)
    for k in range(f.S * f.m):
        a, b = f.tree_offset[k], f.tree_offset[k + 1]
        var, cut, left, right = f.var[a:b], f.cut[a:b], f.left[a:b], f.right[a:b]
        node = np.zeros(N, np.int64)
        for _ in range(b - a):
            v = var[node]
            internal = v >= 0
            if not internal.any():
                break
            vi = np.where(internal, v, 0)
            cv = f.cut_values[f.cut_offsets[vi] + np.where(internal, cut[node], 0)]
            go_right = ~(X[rows, vi] < cv) if f.route == ROUTE_BART else X[rows, vi] > cv
            node = np.where(internal, np.where(go_right, right[node], left[node]), node)
        out[k] = node
    return out


def leaf_ordinals(forest, X):
    """Leaf numbers 0..K-1, leaves counted left-to-right in pre-order (SURVEY App. A7;
    dbarts:::fillPlotCoordinatesForNode numbers them 1..K, BARTInt.R:128-129)."""
    f = as_ns(forest)
    nodes = leaf_nodes(f, X)
    out = np.zeros_like(nodes)
    for k in range(f.S * f.m):
        a, b = f.tree_offset[k], f.tree_offset[k + 1]
        order = preorder(f.left[a:b], f.right[a:b])
        is_leaf = f.var[a:b][order] < 0
        ordinal = np.full(b - a, -1, np.int32)
        ordinal[order[is_leaf]] = np.arange(is_leaf.sum(), dtype=np.int32)
        out[k] = ordinal[nodes[k]]
    return out


def preorder(left, right):
    """Pre-order visiting sequence of relative node indices (root = 0)."""
    order, stack = [], [0]
    while stack:
        k = stack.pop()
        order.append(k)
        if left[k] >= 0:
            stack.append(int(right[k]))
            stack.append(int(left[k]))
    return np.array(order, np.int64)


def tree_sums(forest, X):
    """Sum_t mu_{s,t,leaf(x_i)} in scaled units; (S, N) float64, trees added in order t=0..m-1."""
    f = as_ns(forest)
    nodes = leaf_nodes(f, X)
    N = nodes.shape[1]
    out = np.zeros((f.S, N))
    for s in range(f.S):
        acc = np.zeros(N)
        for t in range(f.m):
            k = t + s * f.m
            acc = acc + f.mu[f.tree_offset[k] + nodes[k]]
        out[s] = acc
    return out


def predict(forest, X):
    """predict.bart [dbarts-mem]: (S, N), original units: (ymax - ymin) * (sum + 0.5) + ymin."""
    f = as_ns(forest)
    return (f.y_max - f.y_min) * (tree_sums(f, X) + 0.5) + f.y_min


# --------------------------------------------------------------------------------------
# a2 / a4 / a9  reductions
# --------------------------------------------------------------------------------------

def col_vars(F):
    """matrixStats::colVars [matrixStats-mem]: per-column unbiased variance (n-1).
    BARTInt.R:314,386; bartMean.R:50.  S == 1 gives NaN like R (0/0)."""
    F = np.asarray(F, np.float64)
    S = F.shape[0]
    mu = (F.astype(LD).sum(axis=0) / LD(S))
    dev = F.astype(LD) - mu
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.asarray((dev * dev).sum(axis=0) / LD(S - 1), np.float64)


def row_means(F):
    """base::rowMeans [R-mem] (long double accumulation). bartMean.R:76,82."""
    F = np.asarray(F, np.float64)
    return np.asarray(F.astype(LD).sum(axis=1) / LD(F.shape[1]), np.float64)


def r_mean(x):
    """base::mean [R-mem]: long double sum / n, then one refinement pass."""
    x = np.asarray(x, np.float64).ravel()
    n = x.size
    s = x.astype(LD).sum() / LD(n)
    s = s + (x.astype(LD) - s).sum() / LD(n)
    return float(s)


def candidate_variance(forest, X, weights=1.0):
    """BARTInt.R:312-314 / bartMean.R:47,50:  colVars(predict(model, X)) * weights."""
    return col_vars(predict(forest, X)) * np.asarray(weights, np.float64)


def finalize_integrals(integrals_scaled, y_min, y_max):
    """BARTInt.R:260-262 (and 290-292): rescale, mean, sd with n-1."""
    integrals = (np.asarray(integrals_scaled, np.float64) + 0.5) * (y_max - y_min) + y_min
    mean = r_mean(integrals)
    n = integrals.size
    with np.errstate(invalid="ignore", divide="ignore"):
        sd = math.sqrt(float(((integrals - mean) ** 2).astype(LD).sum()) / (n - 1)) if n > 1 else float("nan")
    return integrals, mean, sd


def survey_integral(forest, X_full):
    """bartMean.R:74-82: pred = predict(model, fullData); rowMeans; mean(pred);
    sum((rowMeans(pred) - mean)^2) / (nrow(pred) - 1)   (a VARIANCE, quirk B6)."""
    pred = predict(forest, X_full)
    rm = row_means(pred)
    mean_value = r_mean(pred)
    S = pred.shape[0]
    with np.errstate(invalid="ignore", divide="ignore"):
        variance = float(((rm - mean_value) ** 2).astype(LD).sum()) / (S - 1) if S > 1 else float("nan")
    return rm, mean_value, variance


# --------------------------------------------------------------------------------------
# a5 / a7 / a8  exact leaf-probability integral: BARTInt.R:10-223
# --------------------------------------------------------------------------------------

def _pnorm(x):
    # stats::pnorm [R-mem]; erfc form is accurate in both tails
    if x == math.inf:
        return 1.0
    if x == -math.inf:
        return 0.0
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def _pmvnorm1(lower, upper, mean=0.0):
    """mvtnorm::pmvnorm(lower, upper, mean, sigma=1) for length-1 input [mvtnorm-mem]."""
    return _pnorm(upper - mean) - _pnorm(lower - mean)


def _pexp(q):
    """stats::pexp(q) rate 1 [R-mem]."""
    if q <= 0:
        return 0.0
    if q == math.inf:
        return 1.0
    return -math.expm1(-q)


def single_tree_sum(var, cut, left, right, mu, cut_lists, d, measure, lo=None, hi=None):
    """singleTreeSum (BARTInt.R:135-175) on one flattened tree.

    fillProbabilityForNode (:32-72): conditional child probabilities given the current box;
    terminalProbability (:10-30): product of the leaf's and its ancestors' probabilities,
    excluding the root, multiplied leaf-first going up; stump -> probability 1 (:67-69, quirk B12).
    Leaves are summed in data.tree Traverse (pre-order) order (:165-173).
    """
    box = np.empty((2, d))
    if lo is None:
        box[0, :] = 0.0
        box[1, :] = math.inf if measure == "exponential" else 1.0   # :151-155
    else:
        box[0, :] = lo
        box[1, :] = hi
    n = len(var)
    prob = [None] * n
    parent = [-1] * n

    # fillProbabilityForNode, iterative with explicit (node, box) stack; children get their own box copy
    stack = [(0, box)]
    while stack:
        k, cutbox = stack.pop()
        if left[k] >= 0:
            v = int(var[k])
            rule = float(cut_lists[v][int(cut[k])])                      # :45
            lo_v, hi_v = float(cutbox[0, v]), float(cutbox[1, v])
            with np.errstate(all="ignore"):
                if measure == "uniform":                                   # :46-48
                    pl = _div(rule - lo_v, hi_v - lo_v)
                    pr = _div(hi_v - rule, hi_v - lo_v)
                elif measure == "gaussian":                                # :49-52 (quirk B3, as written)
                    norm = _pmvnorm1(lo_v, hi_v, mean=0.5)
                    pl = _div(_pmvnorm1(lo_v, rule, mean=0.0), norm)
                    pr = 1.0 - pl
                elif measure == "gaussian_correct":                        # NOT the reference: both masses about mean 0.5
                    norm = _pmvnorm1(lo_v, hi_v, mean=0.5)
                    pl = _div(_pmvnorm1(lo_v, rule, mean=0.5), norm)
                    pr = _div(_pmvnorm1(rule, hi_v, mean=0.5), norm)
                elif measure == "exponential":                             # :53-56
                    norm = _pexp(hi_v) - _pexp(lo_v)
                    pl = _div(_pexp(rule) - _pexp(lo_v), norm)
                    pr = 1.0 - pl
                else:
                    raise ValueError(f"unknown measure {measure!r}")
            lk, rk = int(left[k]), int(right[k])
            prob[lk], prob[rk] = pl, pr
            parent[lk] = parent[rk] = k
            lbox = cutbox.copy(); lbox[:, v] = (lo_v, rule)                # :59-63
            rbox = cutbox.copy(); rbox[:, v] = (rule, hi_v)                # :64-66
            stack.append((rk, rbox))
            stack.append((lk, lbox))
        elif prob[k] is None:                                              # :67-69
            prob[k] = 1.0

    integral = 0.0
    for k in preorder(left, right):
        if left[k] >= 0:
            continue
        p = prob[k]                                                        # :22
        node = k
        while parent[node] > 0:        # parent is not the root (:24); root has index 0, parent -1
            node = parent[node]
            p = p * prob[node]                                             # :26
        integral = integral + p * float(mu[k])                             # :172
    return integral


def _div(a, b):
    """IEEE division like R (x/0 -> Inf/NaN) without raising."""
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * (math.copysign(1.0, b))


def cut_lists_of(f):
    return [f.cut_values[f.cut_offsets[j]:f.cut_offsets[j + 1]] for j in range(f.d)]


def sample_integrals(forest, measure="uniform", n_draws_used=None, lo=None, hi=None):
    """sampleIntegrals (BARTInt.R:204-223) -> posteriorSum (:177-201) -> singleTreeSum.

    Returns per-draw integrals in SCALED units (caller rescales, :260).  The reference sets
    nDraw <- dim(treeFits)[2] == number of TREES (quirk B1), i.e. only draws 1..m are used;
    pass n_draws_used=min(m, S) for bug-compatible output, None for all S draws.
    posteriorSum uses base::sum (long double accumulation [R-mem]).
    """
    f = as_ns(forest)
    cl = cut_lists_of(f)
    nd = f.S if n_draws_used is None else int(n_draws_used)
    out = np.zeros(nd)
    for s in range(nd):
        acc = LD(0)
        for t in range(f.m):
            k = t + s * f.m
            a, b = f.tree_offset[k], f.tree_offset[k + 1]
            acc += LD(single_tree_sum(f.var[a:b], f.cut[a:b], f.left[a:b], f.right[a:b],
                                      f.mu[a:b], cl, f.d, measure, lo, hi))
        out[s] = float(acc)
    return out


def leaf_probabilities(var, cut, left, right, cut_lists, d, measure="uniform"):
    """P(leaf) for every leaf of one tree (pre-order); helper for the sum-to-one invariant."""
    probs = []
    for k in preorder(left, right):
        if left[k] < 0:
            mu = np.zeros(len(var)); mu[k] = 1.0
            probs.append(single_tree_sum(var, cut, left, right, mu, cut_lists, d, measure))
    return np.array(probs)


# --------------------------------------------------------------------------------------
# BART::wbart / pbart tree draws (reference README.md:202 invites swapping BART packages)
# --------------------------------------------------------------------------------------

def forest_from_wbart(trees_text, cut_lists, offset=0.0):
    """Flat forest from BART::wbart's ``fit$treedraws`` [BART-mem].

    ``trees_text`` = treedraws$trees: first line "ndpost ntree p"; then, draw by draw and tree by tree, a line with
    the node count followed by one line "nid var cut theta" per node (tree.cpp operator<<).  nid is the heap index
    (root 1, children 2*nid and 2*nid+1); a node is internal iff its left child is listed; var and cut are 0-based,
    cut indexes ``treedraws$cutpoints[[var+1]]``; theta is the leaf value (ignored on internal nodes).
    Routing (tree::bn): left iff x[var] < cutpoints[var][cut].  predict.wbart = offset + sum_t theta with
    offset = fit$offset = mean(y.train); expressed with the dbarts rescale (ymax - ymin) * (sum + 0.5) + ymin by
    setting ymin = offset - 0.5, ymax = offset + 0.5.
    """
    tok = trees_text.split()
    S, m, d = int(tok[0]), int(tok[1]), int(tok[2])
    if d != len(cut_lists):
        raise ValueError("treedraws header says p = %d but %d cut-point lists were given" % (d, len(cut_lists)))
    pos = 3
    tree_offset = [0]
    var, cut, left, right, mu = [], [], [], [], []
    for _ in range(S * m):
        nn = int(tok[pos]); pos += 1
        nodes = {}
        for _k in range(nn):
            nid, v, c, th = int(tok[pos]), int(tok[pos + 1]), int(tok[pos + 2]), float(tok[pos + 3])
            pos += 4
            nodes[nid] = (v, c, th)
        if 1 not in nodes:
            raise ValueError("tree without a root node")
        base = len(var)
        stack = [(1, -1, False)]          # (nid, parent slot, is right child): pre-order, left child first
        while stack:
            nid, parent, is_right = stack.pop()
            slot = len(var) - base
            if parent >= 0:
                (right if is_right else left)[base + parent] = slot
            v, c, th = nodes[nid]
            if 2 * nid in nodes:
                if 2 * nid + 1 not in nodes:
                    raise ValueError("node %d has a left child but no right child" % nid)
                if not (0 <= v < d) or not (0 <= c < len(cut_lists[v])):
                    raise ValueError("split (%d, %d) outside the cut-point lists" % (v, c))
                var.append(v); cut.append(c); left.append(-1); right.append(-1); mu.append(0.0)
                stack.append((2 * nid + 1, slot, True))
                stack.append((2 * nid, slot, False))
            else:
                var.append(-1); cut.append(0); left.append(-1); right.append(-1); mu.append(th)
        if len(var) - base != nn:
            raise ValueError("unreachable nodes in a tree")
        tree_offset.append(len(var))
    if pos != len(tok):
        raise ValueError("trailing tokens after the last tree")
    cut_offsets = np.zeros(d + 1, np.int32)
    for j in range(d):
        cut_offsets[j + 1] = cut_offsets[j] + len(cut_lists[j])
    cut_values = np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0)
    return SimpleNamespace(S=S, m=m, d=d, tree_offset=np.array(tree_offset, np.int64), var=np.array(var, np.int32),
                           cut=np.array(cut, np.int32), left=np.array(left, np.int32), right=np.array(right, np.int32),
                           mu=np.array(mu, np.float64), cut_offsets=cut_offsets, cut_values=cut_values,
                           y_min=float(offset) - 0.5, y_max=float(offset) + 0.5, route=ROUTE_BART)


def forest_from_value_splits(S, m, d, tree_offset, var, value, left=None, right=None, route=ROUTE_DBARTS, leaf_scale=1.0,
                             offset=0.0):
    """Flat forest from trees whose splits are VALUES [dbarts-mem / bartMachine-mem] -- what newer BART packages hand
    out instead of the dbarts 0.9-8 tree strings that getTree() (src/BARTInt.R:92-133) parses:

      dbarts >= 0.9-9   ``fit$getTrees()``: data frame, one row per node in pre-order, columns sample, tree, n,
                        var (-1 = leaf), value (split value | leaf prediction);
      bartMachine       ``extract_raw_node_data``: nested nodes (splitAttributeM, splitValue, y_pred, left, right).

    A point goes left iff x <= split value (route 0; dbarts and bartMachine) or x < split value (route 1).  The cut
    lists are, per variable, the sorted distinct split values, so "x > value" <=> "code > index of value".
    predict = offset + leaf_scale * sum_t leaf value, written with the dbarts rescale by ymin = offset - 0.5.
    left/right None: nodes are in pre-order (left child = next node, right child after the left subtree)."""
    tree_offset = np.asarray(tree_offset, np.int64)
    var = np.asarray(var, np.int32)
    value = np.asarray(value, np.float64)
    nn = int(tree_offset[-1])
    cut_lists = [np.unique(value[(var == j)]) for j in range(d)]
    if left is None:
        left = np.full(nn, -1, np.int32)
        right = np.full(nn, -1, np.int32)
        for k in range(S * m):
            b, e = int(tree_offset[k]), int(tree_offset[k + 1])
            pend = []
            for a in range(e - b):
                if var[b + a] >= 0:
                    left[b + a] = a + 1
                    pend.append(a)
                elif pend:
                    right[b + pend.pop()] = a + 1
            if pend or any(right[b + a] >= e - b or left[b + a] >= e - b for a in range(e - b)):
                raise ValueError("tree %d is not a complete binary tree in pre-order" % k)
    cut = np.zeros(nn, np.int32)
    mu = np.zeros(nn)
    for g in range(nn):
        if var[g] >= 0:
            cut[g] = int(np.searchsorted(cut_lists[var[g]], value[g]))
        else:
            mu[g] = leaf_scale * value[g]
    cut_offsets = np.zeros(d + 1, np.int32)
    cut_offsets[1:] = np.cumsum([len(c) for c in cut_lists])
    cut_values = np.concatenate(cut_lists) if d else np.zeros(0)
    return SimpleNamespace(S=S, m=m, d=d, tree_offset=tree_offset, var=var, cut=cut, left=np.asarray(left, np.int32),
                           right=np.asarray(right, np.int32), mu=mu, cut_offsets=cut_offsets, cut_values=cut_values,
                           y_min=float(offset) - 0.5, y_max=float(offset) + 0.5, route=int(route))


def wbart_trees_text(forest) -> str:
    """Inverse of forest_from_wbart: writes a flat forest in the treedraws$trees text format (heap node ids, nodes in
    pre-order as tree::getnodes lists them; theta of internal nodes written as 0)."""
    f = as_ns(forest)
    out = ["%d %d %d" % (f.S, f.m, f.d)]
    for k in range(f.S * f.m):
        a, b = int(f.tree_offset[k]), int(f.tree_offset[k + 1])
        out.append(str(b - a))
        stack = [(0, 1)]
        while stack:
            node, nid = stack.pop()
            if f.var[a + node] >= 0:
                out.append("%d %d %d 0" % (nid, f.var[a + node], f.cut[a + node]))
                stack.append((int(f.right[a + node]), 2 * nid + 1))
                stack.append((int(f.left[a + node]), 2 * nid))
            else:
                out.append("%d 0 0 %r" % (nid, float(f.mu[a + node])))
    return "\n".join(out) + "\n"


# --------------------------------------------------------------------------------------
# 8f item 2: on-device sampling of the integration measure -- restatement of the generator
# (Philox4x32-10, Salmon et al. 2011; the measure transforms of BARTInt.R:301,304,307)
# --------------------------------------------------------------------------------------

def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 on uint32 arrays; returns the first two output words."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32)


def generate_points(measure, N, d, seed, first_index=0):
    """(N, d) float64 points exactly as generate_points_kernel draws them (uniform: bit-exact; the gaussian and
    exponential transforms go through libm / scipy and may differ from the device in the last ulp)."""
    i = (np.arange(N, dtype=np.uint64) + np.uint64(first_index))[:, None]
    j = np.arange(d, dtype=np.uint64)[None, :]
    w0, w1 = philox4x32_10(i & np.uint64(0xFFFFFFFF), i >> np.uint64(32), j, np.uint64(0),
                           seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = ((w0 >> np.uint32(5)).astype(np.float64) * 67108864.0 + (w1 >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0
    if measure == "uniform":
        return u
    if measure == "exponential":
        return -np.log1p(-u)
    from scipy.special import ndtri
    plo, phi = 0.30853753872598690, 0.69146246127401310
    return 0.5 + ndtri(plo + (phi - plo) * u)
