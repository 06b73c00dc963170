"""CPU tests of the oracle itself: pin the numpy/C restatements against the hand-derived KATs and the
invariants of the reference's estimators (SURVEY.md section 8c "substitute pins")."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle_np as onp
from bartint_testlib import kat_forest

from bartint_b200 import synth


def test_kats_oracle(kats):
    for k in kats:
        ff = kat_forest(k)
        X = np.array(k["X"], dtype=float)
        assert np.array_equal(onp.leaf_ordinals(ff, X), np.array(k["leaf"])), k["name"]
        np.testing.assert_allclose(onp.tree_sums(ff, X), np.array(k["tree_sum"]), rtol=0, atol=1e-15, err_msg=k["name"])
        np.testing.assert_allclose(onp.predict(ff, X), np.array(k["predict"]), rtol=1e-15, atol=1e-15, err_msg=k["name"])
        for measure, want in k["integrals"].items():
            got = onp.sample_integrals(ff, measure)
            np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-15, err_msg=f"{k['name']} {measure}")
        if "col_vars" in k:
            np.testing.assert_allclose(onp.col_vars(onp.predict(ff, X)), k["col_vars"], rtol=1e-13)
            np.testing.assert_allclose(onp.row_means(onp.predict(ff, X)), k["row_means"], rtol=1e-14)


def test_finalize_matches_reference_formulas():
    # BARTInt.R:260-262 on a hand example: integrals (scaled) -0.5, 0, 0.5 with y in [1, 3] -> 1, 2, 3
    res, mean, sd = onp.finalize_integrals([-0.5, 0.0, 0.5], 1.0, 3.0)
    np.testing.assert_allclose(res, [1.0, 2.0, 3.0])
    assert mean == 2.0 and sd == 1.0
    _, _, sd1 = onp.finalize_integrals([0.1], 0.0, 1.0)
    assert math.isnan(sd1)          # sd of one draw is NaN in R


def test_col_vars_is_unbiased_variance():
    rng = np.random.default_rng(0)
    F = rng.normal(size=(17, 5))
    np.testing.assert_allclose(onp.col_vars(F), F.var(axis=0, ddof=1), rtol=1e-13)
    assert np.isnan(onp.col_vars(F[:1])).all()      # one draw: 0/0 like matrixStats


def test_bin_codes_equal_routing(small_case):
    ff, _, _, X = small_case
    Xe = X.copy()
    cl = ff.cut_lists()
    # plant exact ties, nextafter, out of range and NaN values
    Xe[0, 0] = cl[0][10]; Xe[1, 0] = np.nextafter(cl[0][10], np.inf); Xe[2, 1] = -5.0; Xe[3, 1] = 9.0
    Xe[4, 2] = np.nan; Xe[5, 3] = cl[3][0]; Xe[6, 3] = cl[3][-1]; Xe[7, 3] = np.nextafter(cl[3][-1], np.inf)
    codes = onp.bin_points(Xe, ff.cut_offsets, ff.cut_values)
    assert codes[0, 0] == 10 and codes[1, 0] == 11 and codes[2, 1] == 0 and codes[3, 1] == 100
    assert codes[4, 2] == 0 and codes[5, 3] == 0 and codes[6, 3] == 99 and codes[7, 3] == 100
    # "right iff code > cut index" must reproduce "right iff x > cut value" for every node
    leaves = onp.leaf_nodes(ff, Xe)
    f = onp.as_ns(ff)
    for k in range(0, f.S * f.m, 7):
        a, b = f.tree_offset[k], f.tree_offset[k + 1]
        node = np.zeros(len(Xe), np.int64)
        for _ in range(b - a):
            v = f.var[a:b][node]; internal = v >= 0
            if not internal.any():
                break
            vi = np.where(internal, v, 0)
            go = codes[np.arange(len(Xe)), vi] > f.cut[a:b][node]
            node = np.where(internal, np.where(go, f.right[a:b][node], f.left[a:b][node]), node)
        assert np.array_equal(node, leaves[k])


def test_leaf_probabilities_sum_to_one(small_case):
    ff, *_ = small_case
    f = onp.as_ns(ff)
    cl = onp.cut_lists_of(f)
    for k in range(f.S * f.m):
        a, b = f.tree_offset[k], f.tree_offset[k + 1]
        p = onp.leaf_probabilities(f.var[a:b], f.cut[a:b], f.left[a:b], f.right[a:b], cl, f.d, "uniform")
        assert abs(p.sum() - 1.0) < 1e-13
        # the synthetic cut points lie inside the training range, a subset of [0,1] -> all volumes in [0,1]
        assert (p >= 0).all()


def test_exact_integral_equals_aligned_grid_quadrature():
    # d = 2: integrate the piecewise-constant ensemble exactly on the grid induced by all cut points
    rng = np.random.default_rng(3)
    x_train = rng.random((25, 2))
    ff = synth.prior_forest(9, 6, 5, x_train, numcut=12)
    exact = onp.sample_integrals(ff, "uniform")
    edges = [np.unique(np.concatenate([[0.0, 1.0], np.clip(c, 0, 1)])) for c in ff.cut_lists()]
    mids = [0.5 * (e[1:] + e[:-1]) for e in edges]
    w = np.outer(np.diff(edges[0]), np.diff(edges[1])).ravel()
    G = np.array([[a, b] for a in mids[0] for b in mids[1]])
    quad = onp.tree_sums(ff, G) @ w
    np.testing.assert_allclose(exact, quad, rtol=0, atol=2e-15 * ff.m)


def test_mc_estimate_within_5_sigma(small_case):
    ff, _, _, X = small_case
    F = onp.tree_sums(ff, X)
    exact = onp.sample_integrals(ff, "uniform")
    se = F.std(axis=1, ddof=1) / math.sqrt(X.shape[1] if False else X.shape[0])
    assert (np.abs(F.mean(axis=1) - exact) <= 5 * se + 1e-12).all()


def test_dbarts_string_round_trip(small_case):
    ff, x_train, *_ = small_case
    strings, fits = ff.to_dbarts_state(x_train)
    assert strings.shape == (ff.m, ff.S) and fits.shape == (len(x_train), ff.m, ff.S)
    back = onp.forest_from_dbarts(strings, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max)
    for key in ("tree_offset", "var", "cut", "left", "right", "mu"):
        assert np.array_equal(np.asarray(back[key]), getattr(ff, key)), key
    assert onp.tokenize_tree_string("0 49 .1 20 ..") == ["0", "49", ".", "1", "20", ".", "."]
    with pytest.raises(ValueError):
        onp.parse_tree_tokens(onp.tokenize_tree_string("0 49 ."))


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.integers(1, 4), st.integers(1, 4))
def test_survey_integral_consistency(seed, d, S, m):
    rng = np.random.default_rng(seed)
    x_train = rng.random((12, d))
    ff = synth.prior_forest(seed, S, m, x_train, rng.normal(size=12), numcut=7)
    X = rng.random((50, d))
    rm, mean_value, variance = onp.survey_integral(ff, X)
    P = onp.predict(ff, X)
    np.testing.assert_allclose(rm, P.mean(axis=1), rtol=1e-13, atol=1e-13)
    assert abs(mean_value - P.mean()) < 1e-12
    if S > 1:
        assert abs(variance - P.mean(axis=1).var(ddof=1)) < 1e-12
    else:
        assert math.isnan(variance)


def test_corrected_gaussian_measure_is_a_probability_and_matches_mc(small_case):
    """'gaussian_correct' (not the reference's as-written branch): leaf masses sum to one and the
    exact integral agrees with Monte Carlo under N(0.5, 1) truncated to [0, 1]^d."""
    ff, *_ = small_case
    f = onp.as_ns(ff)
    cl = onp.cut_lists_of(f)
    for k in range(0, f.S * f.m, 5):
        a, b = f.tree_offset[k], f.tree_offset[k + 1]
        p = onp.leaf_probabilities(f.var[a:b], f.cut[a:b], f.left[a:b], f.right[a:b], cl, f.d, "gaussian_correct")
        assert abs(p.sum() - 1.0) < 1e-12 and (p >= -1e-15).all()
    rng = np.random.default_rng(3)
    X = synth.sample_measure(rng, "gaussian", 40000, ff.d)
    F = onp.tree_sums(ff, X)
    exact = onp.sample_integrals(ff, "gaussian_correct")
    se = F.std(axis=1, ddof=1) / math.sqrt(X.shape[0])
    assert (np.abs(F.mean(axis=1) - exact) <= 5 * se + 1e-12).all()
    from oracle import oracle_c as oc
    np.testing.assert_allclose(oc.exact_integrals(ff, "gaussian_correct"), exact, rtol=1e-13, atol=1e-15)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (first two output words) pin the generator's restatement."""
    u = np.uint64
    for ctr, key, want in [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d)),
                           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e)),
                           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
                            (0xd16cfe09, 0x94fdcceb))]:
        a, b = onp.philox4x32_10(u(ctr[0]), u(ctr[1]), u(ctr[2]), u(ctr[3]), key[0], key[1])
        assert (int(a), int(b)) == want
