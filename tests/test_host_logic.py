"""CPU tests of the host-side logic: synthetic generator, wire format, R-semantics helpers."""
import numpy as np

from bartint_b200 import api, synth


def test_prior_forest_shape_and_determinism():
    rng = np.random.default_rng(4)
    x = rng.random((60, 5))
    a = synth.prior_forest(3, 40, 20, x)
    b = synth.prior_forest(3, 40, 20, x)
    for k in ("tree_offset", "var", "cut", "left", "right", "mu"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    nodes = np.diff(a.tree_offset)
    assert (nodes % 2 == 1).all()
    leaves = (nodes + 1) // 2
    assert 1.8 < leaves.mean() < 3.2          # BART prior: E[leaves] ~ 2.5
    # pre-order: left child is always the next node
    internal = a.var >= 0
    first = np.repeat(a.tree_offset[:-1], nodes)
    assert (a.left[internal] == (np.arange(a.n_nodes) - first + 1)[internal]).all()
    assert (a.cut[internal] >= 0).all() and (a.cut[internal] < 100).all()


def test_every_leaf_owns_a_training_row():
    rng = np.random.default_rng(5)
    x = rng.random((30, 3))
    ff = synth.prior_forest(8, 10, 6, x)
    _, fits = ff.to_dbarts_state(x)
    # each leaf value of each tree appears among the fits of that tree
    for s in range(ff.S):
        for t in range(ff.m):
            k = t + s * ff.m
            a, b = ff.tree_offset[k], ff.tree_offset[k + 1]
            leaf_mu = ff.mu[a:b][ff.var[a:b] < 0]
            assert set(np.round(leaf_mu, 15)) == set(np.round(fits[:, t, s], 15))


def test_tree_string_grammar():
    ff = synth.forest_from_trees([(0, 49, 0.1, (1, 20, 0.2, 0.3))], 1, 1, [np.linspace(0, 1, 102)[1:-1]] * 2)
    assert ff.tree_string(0) == "0 49 .1 20 .."          # the example of SURVEY App. A2
    ff = synth.forest_from_trees([0.7], 1, 1, [np.array([0.5])])
    assert ff.tree_string(0) == "."


def test_uniform_cutpoints_formula():
    x = np.array([[0.0], [1.0], [0.25]])
    c = synth.uniform_cutpoints(x, numcut=3)[0]
    np.testing.assert_allclose(c, [0.25, 0.5, 0.75])      # min + (k+1)(max-min)/(numcut+1)


def test_r_sample_one_scalar_quirk():
    rng = np.random.default_rng(0)
    # unique argmax at 1-based index 7 -> R's sample(7, 1) draws from 1:7 (quirk B2)
    draws = {api.r_sample_one(np.array([7]), rng) for _ in range(300)}
    assert draws == set(range(1, 8))
    # ties -> one of the tied indices
    draws = {api.r_sample_one(np.array([3, 9]), rng) for _ in range(100)}
    assert draws == {3, 9}


def test_adversarial_generators_are_valid_trees():
    for kind in ("stumps", "left_chain", "right_chain", "same_var", "balanced", "edge_cuts"):
        ff = synth.adversarial_forest(kind, 3, 4, 3, numcut=20, seed=1, depth=9)
        nodes = np.diff(ff.tree_offset)
        assert (nodes % 2 == 1).all()
        assert (ff.var < 3).all() and (ff.cut < 20).all()


def test_wire_strings_slice_is_a_view():
    import ctypes as C
    from bartint_b200 import WireStrings
    strings = np.array([[f"t{t}s{s}" for s in range(5)] for t in range(3)], dtype=object)      # [m=3, S=5]
    w = WireStrings(strings)
    v = w.slice_draws(1, 4)
    assert (v.m, v.S) == (3, 3)
    assert [v.array[k] for k in range(9)] == [f"t{t}s{s}".encode() for s in range(1, 4) for t in range(3)]
    assert C.addressof(v.array) == C.addressof(w.array) + 3 * C.sizeof(C.c_char_p)


def test_candidate_pick_follows_r_recycling(monkeypatch):
    """BARTInt.R:305,308,314-315: `colVars(fValues) * weights` with weights that are not length C.
    (R's sample(scalar) quirk B2, which _pick keeps, is switched off here so that the recycling alone is visible.)"""
    from bartint_b200 import api
    monkeypatch.setattr(api, "r_sample_one", lambda v, rng: int(np.asarray(v)[0]))
    rng = np.random.default_rng(0)
    C_, d = 100, 3
    var = rng.random(C_)
    # exponential: length-d weights recycled over the candidates -> weight of candidate i is w[i mod d]
    w = np.array([1.0, 5.0, 0.1])
    want = int(np.argmax(var * np.array([w[i % d] for i in range(C_)]))) + 1
    assert api._pick(var, w, np.random.default_rng(1)) == want
    # gaussian: C x d matrix; R's product is C x d with var down the columns, which() gives column-major linear indices
    W = rng.random((C_, d))
    prod = np.array([[var[i] * W[i, j] for j in range(d)] for i in range(C_)])
    lin = int(np.argmax(prod.ravel(order="F")))            # 0-based linear index
    assert api._pick(var, W, np.random.default_rng(1)) == lin % C_ + 1
    # scalar / length-C weights: the plain product
    assert api._pick(var, np.ones(1), np.random.default_rng(1)) == int(np.argmax(var)) + 1
    wc = rng.random(C_)
    assert api._pick(var, wc, np.random.default_rng(1)) == int(np.argmax(var * wc)) + 1
    # the weights objects _candidates builds have the reference's shapes
    X, wg = api._candidates("gaussian", d, C_, np.random.default_rng(2))
    assert wg.shape == (C_, d)
    X, we = api._candidates("exponential", d, C_, np.random.default_rng(2))
    assert we.shape == (d,) and np.allclose(we, np.exp(-X).mean(axis=0))
    assert api._candidates("exponential", d, C_, np.random.default_rng(2), refcompat=False)[1].shape == (C_,)
