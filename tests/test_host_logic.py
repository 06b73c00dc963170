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
