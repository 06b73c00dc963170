"""Exporter for trees given with split VALUES (include/bartint.h: bartint_forest_from_value_splits) -- the format of
dbarts >= 0.9-9 (`fit$getTrees()`: pre-order rows sample / tree / var / value) and of bartMachine's raw node data, which
replace the dbarts 0.9-8 tree strings that getTree() (src/BARTInt.R:92-133) parses (README.md:102-106, :202).
CPU: the host exporter against the oracle's restatement and against a direct recursive evaluation on the values.
GPU: predictions / draw means / variances of the uploaded forest against the same direct evaluation."""
import numpy as np
import pytest

import hosthooks
from bartint_b200 import BartIntError, FlatForest, Forest, synth
from oracle import oracle_np as onp


def _random_value_forest(seed, S, m, d, explicit_children, shuffled=False):
    """Trees with real-valued splits (a few values per variable so that values repeat across trees), node arrays either in
    pre-order (children implied) or with explicit child indices in a shuffled node order."""
    rng = np.random.default_rng(seed)
    grid = [np.sort(rng.random(rng.integers(3, 40))) for _ in range(d)]
    off, var, val, left, right = [0], [], [], [], []
    nested = []

    def grow(depth):
        if depth >= 6 or rng.random() > 0.75 / (1 + depth) ** 0.7:
            return float(rng.normal(0, 0.1))
        v = int(rng.integers(d))
        return (v, float(rng.choice(grid[v])), grow(depth + 1), grow(depth + 1))

    for _ in range(S * m):
        tree = grow(0)
        nested.append(tree)
        nodes = []

        def rec(node):
            me = len(nodes)
            if isinstance(node, tuple):
                nodes.append([node[0], node[1], -1, -1])
                nodes[me][2] = rec(node[2])
                nodes[me][3] = rec(node[3])
            else:
                nodes.append([-1, node, -1, -1])
            return me
        rec(tree)
        order = np.arange(len(nodes))
        if shuffled and len(nodes) > 1:           # any node order is fine with explicit children, as long as the root is node 0
            order = np.concatenate([[0], 1 + rng.permutation(len(nodes) - 1)])
        pos = np.empty(len(nodes), int)
        pos[order] = np.arange(len(nodes))
        for a in order:
            v, x, l, r = nodes[a]
            var.append(v); val.append(x)
            left.append(-1 if l < 0 else int(pos[l])); right.append(-1 if r < 0 else int(pos[r]))
        off.append(len(var))
    kw = dict(left=np.array(left, np.int32), right=np.array(right, np.int32)) if explicit_children else {}
    return nested, (S, m, d, np.array(off, np.int64), np.array(var, np.int32), np.array(val)), kw


def _direct(nested, S, m, X, route, leaf_scale, offset):
    """predict() by walking the nested trees on the VALUES (no cut lists, no oracle code)."""
    out = np.zeros((S, len(X)))
    for k, tree in enumerate(nested):
        for i, x in enumerate(X):
            node = tree
            while isinstance(node, tuple):
                v, c, l, r = node
                go_left = (x[v] < c) if route == 1 else (x[v] <= c)
                node = l if go_left else r
            out[k // m, i] += leaf_scale * node
    return out + offset


@pytest.mark.parametrize("explicit,shuffled,route", [(False, False, 0), (True, False, 0), (True, True, 1), (False, False, 1)])
def test_exporter_matches_oracle_and_direct_walk(explicit, shuffled, route):
    nested, args, kw = _random_value_forest(5 + route, 6, 7, 4, explicit, shuffled)
    S, m, d, off, var, val = args
    got = hosthooks.export_value_splits(S, m, d, off, var, val, leaf_scale=2.5, **kw)
    want = onp.forest_from_value_splits(S, m, d, off, var, val, kw.get("left"), kw.get("right"), route, 2.5, 0.75)
    assert np.array_equal(got["cut_offsets"], want.cut_offsets) and np.array_equal(got["cut_values"], want.cut_values)
    ff = FlatForest(S, m, d, got["tree_offset"], got["var"], got["cut"], got["left"], got["right"], got["mu"],
                    got["cut_offsets"], got["cut_values"], 0.25, 1.25, route=route)
    rng = np.random.default_rng(1)
    X = rng.random((300, d))
    X[::5, 1] = rng.choice(got["cut_values"][got["cut_offsets"][1]:got["cut_offsets"][2]], len(X[::5, 1]))   # on split values
    direct = _direct(nested, S, m, X, route, 2.5, 0.75)
    np.testing.assert_allclose(onp.predict(ff, X), direct, rtol=0, atol=1e-13)          # exporter's SoA, oracle's walk
    np.testing.assert_allclose(onp.predict(want, X), direct, rtol=0, atol=1e-13)        # oracle's restatement end to end


def test_rejects_malformed_input():
    off = np.array([0, 3], np.int64)
    with pytest.raises(BartIntError):        # pre-order that stops in the middle of a tree
        hosthooks.export_value_splits(1, 1, 2, off, np.array([0, 1, -1], np.int32), np.array([0.5, 0.2, 1.0]))
    with pytest.raises(BartIntError):        # variable outside the model
        hosthooks.export_value_splits(1, 1, 2, off, np.array([5, -1, -1], np.int32), np.array([0.5, 0.2, 1.0]))
    with pytest.raises(BartIntError):        # NaN split value
        hosthooks.export_value_splits(1, 1, 2, off, np.array([0, -1, -1], np.int32), np.array([np.nan, 0.2, 1.0]))
    # more than 255 distinct split values of one variable do not fit uint8 codes
    n = 300
    var = np.array(([0, -1] * n) + [-1], np.int32)         # a right-leaning chain in pre-order: split, leaf, split, leaf, ...
    var = np.concatenate([np.tile([0], 1), var[1:]])
    chain_var, chain_val = [], []
    for k in range(n):
        chain_var += [0, -1]; chain_val += [k / n, 0.0]
    chain_var.append(-1); chain_val.append(1.0)
    with pytest.raises(BartIntError) as e:
        hosthooks.export_value_splits(1, 1, 1, np.array([0, len(chain_var)], np.int64), np.array(chain_var, np.int32), np.array(chain_val))
    assert e.value.status == 4


@pytest.mark.gpu
@pytest.mark.parametrize("route", [0, 1])
def test_gpu_predictions_from_value_splits(route):
    nested, args, kw = _random_value_forest(11 + route, 40, 9, 5, explicit_children=bool(route))
    S, m, d, off, var, val = args
    f = Forest.from_value_splits(S, m, d, off, var, val, route=route, leaf_scale=0.5, offset=3.0, **kw)
    rng = np.random.default_rng(2)
    X = rng.random((700, d))
    cl = f.cutpoints()
    for j in range(d):
        if len(cl[j]):
            X[j::9, j] = rng.choice(cl[j], len(X[j::9, j]))
    direct = _direct(nested, S, m, X, route, 0.5, 3.0)
    out = f.eval(X, want=("full_pred", "draw_mean", "point_var"))
    np.testing.assert_allclose(out["full_pred"], direct, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["draw_mean"], direct.mean(axis=1), rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["point_var"], direct.var(axis=0, ddof=1), rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(f.eval(X, want=("draw_mean",))["draw_mean"], direct.mean(axis=1), rtol=0, atol=1e-12)   # bit-sliced
    f.close()
