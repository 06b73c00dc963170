"""BART::wbart tree-draw exporter: text format, routing rule "left iff x < cut", offset.

CPU part: the oracle restatement against hand-computed known answers and format round trips.
GPU part (-m gpu): bartint_forest_from_wbart through the C ABI against the oracle -- cut codes and leaf indices
bit-exact (including points ON a cut value and NaN, where the rule differs from dbarts'), predictions and reductions
within 1e-6 (asserted much tighter), exact integrals identical to the same trees under the dbarts rule (the two rules
differ on a set of measure zero).
"""
import numpy as np
import pytest

from oracle import oracle_np as onp
from bartint_b200 import FlatForest, Forest, synth, BartIntError

# one draw, two trees, two variables; cut points x0: .25 .5 .75, x1: .5
#   tree 0: root (x0, cut index 1 = .5): left leaf -0.25, right -> (x1, cut 0 = .5): left 0.5, right 1.5
#   tree 1: stump 0.125
KAT_TEXT = "1 2 2\n5\n1 0 1 0\n2 0 0 -0.25\n3 1 0 0\n6 0 0 0.5\n7 0 0 1.5\n1\n1 0 0 0.125\n"
KAT_CUTS = [[0.25, 0.5, 0.75], [0.5]]
KAT_OFFSET = 10.0
KAT_X = np.array([[0.49, 0.9],                       # x0 < .5 -> left leaf            : 10 - .25 + .125
                  [0.5, 0.2],                        # x0 == .5 -> RIGHT (dbarts: left); x1 < .5 -> 0.5
                  [0.5, 0.5],                        # x1 == .5 -> right               : 1.5
                  [np.nan, np.nan],                  # NaN: "x < c" false -> right, right
                  [0.75, 0.49999]])
KAT_PRED = np.array([[9.875, 10.625, 11.625, 11.625, 10.625]])
KAT_LEAF = np.array([[0, 1, 2, 2, 1], [0, 0, 0, 0, 0]])
KAT_CODES = np.array([[1, 1], [2, 0], [2, 1], [3, 1], [3, 0]])


def test_oracle_wbart_known_answers():
    f = onp.forest_from_wbart(KAT_TEXT, KAT_CUTS, KAT_OFFSET)
    assert (f.S, f.m, f.d, f.route) == (1, 2, 2, onp.ROUTE_BART)
    assert f.tree_offset.tolist() == [0, 5, 6]
    assert f.var.tolist() == [0, -1, 1, -1, -1, -1] and f.cut[[0, 2]].tolist() == [1, 0]
    assert f.left[[0, 2]].tolist() == [1, 3] and f.right[[0, 2]].tolist() == [2, 4]
    assert np.array_equal(onp.predict(f, KAT_X), KAT_PRED)
    assert np.array_equal(onp.leaf_ordinals(f, KAT_X), KAT_LEAF)
    assert np.array_equal(onp.bin_points(KAT_X, f.cut_offsets, f.cut_values, route=onp.ROUTE_BART), KAT_CODES)
    # the same trees under the dbarts rule route x == cut and NaN to the LEFT
    f.route = onp.ROUTE_DBARTS
    assert np.array_equal(onp.predict(f, KAT_X), np.array([[9.875, 9.875, 9.875, 9.875, 10.625]]))


def test_oracle_wbart_codes_decide_routing():
    """right iff code > cut index must hold under both rules (that is all the evaluation kernels rely on)."""
    rng = np.random.default_rng(5)
    cuts = np.sort(rng.random(17))
    x = np.concatenate([cuts, np.nextafter(cuts, np.inf), np.nextafter(cuts, -np.inf), rng.random(50), [-1.0, 2.0, np.nan]])
    off, vals = np.array([0, len(cuts)], np.int32), cuts
    for route in (onp.ROUTE_DBARTS, onp.ROUTE_BART):
        codes = onp.bin_points(x[:, None], off, vals, route=route)[:, 0]
        for c in range(len(cuts)):
            with np.errstate(invalid="ignore"):
                right = ~(x < cuts[c]) if route == onp.ROUTE_BART else (x > cuts[c])
            assert np.array_equal(codes > c, right), (route, c)


def test_wbart_text_round_trip_and_c_oracle():
    from oracle import oracle_c as oc
    rng = np.random.default_rng(11)
    x_train = rng.random((60, 4))
    ff = synth.prior_forest(11, 6, 9, x_train, None)
    text = ff.to_wbart_text()
    g = onp.forest_from_wbart(text, ff.cut_lists(), offset=-3.0)
    for k in ("tree_offset", "var", "cut", "left", "right", "mu"):
        assert np.array_equal(getattr(g, k), getattr(ff, k)), k
    assert onp.wbart_trees_text(g) == text
    X = rng.random((300, 4))
    X[:40, 0] = rng.choice(ff.cut_lists()[0], 40)           # points exactly on cut values
    X[40:50, 1] = np.nan
    P = onp.predict(g, X)
    out = oc.predict_reduce(g, X, nthreads=2, want=("full_pred", "draw_mean"))
    assert np.allclose(out["full_pred"], P, rtol=0, atol=1e-13)
    assert np.array_equal(oc.leaf_ordinals(g, X), onp.leaf_ordinals(g, X))
    # and the rule matters on this input
    g.route = onp.ROUTE_DBARTS
    assert not np.array_equal(onp.leaf_ordinals(g, X), oc.leaf_ordinals(onp.forest_from_wbart(text, ff.cut_lists()), X))


@pytest.mark.parametrize("bad", [
    "", "1 1\n1\n1 0 0 0.5\n", "1 1 2\n2\n1 0 0 0\n2 0 0 1.0\n",            # header; missing right child
    "1 1 2\n3\n1 5 0 0\n2 0 0 1\n3 0 0 2\n", "1 1 2\n3\n1 0 9 0\n2 0 0 1\n3 0 0 2\n",   # variable / cut out of range
    "1 1 2\n1\n2 0 0 1.0\n", "1 1 2\n2\n1 0 0 1.0\n4 0 0 1.0\n",            # no root; unreachable node
])
def test_oracle_wbart_rejects_malformed(bad):
    with pytest.raises((ValueError, IndexError)):
        onp.forest_from_wbart(bad, KAT_CUTS, 0.0)


# ------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["gather", "table"])
def kernel_kind(request, monkeypatch):
    monkeypatch.setenv("BARTINT_KERNEL", request.param)
    return request.param


def _close(got, want, tol, scale):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= tol * scale, np.max(np.abs(got - want)) / scale


@pytest.mark.gpu
def test_wbart_kat_cuda(kernel_kind):
    f = Forest.from_wbart(KAT_TEXT, KAT_CUTS, KAT_OFFSET)
    assert (f.S, f.m, f.d, f.route) == (1, 2, 2, 1)
    pts = f.points(KAT_X)
    assert np.array_equal(pts.codes().astype(np.int64), KAT_CODES)
    pts.close()
    for table in (False, True):
        assert np.array_equal(f.leaf_indices(KAT_X, use_table_kernel=table), KAT_LEAF)
    for walk in (False, True):
        out = f.eval(KAT_X, want=("full_pred", "draw_mean"), force_walk=walk)
        assert np.array_equal(out["full_pred"], KAT_PRED)
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("d", [3, 10, 20])
def test_wbart_forest_matches_oracle(d, kernel_kind):
    rng = np.random.default_rng(100 + d)
    x_train = rng.random((20 * d, d))
    ff = synth.prior_forest(7 + d, 12, 25, x_train, None)
    text = ff.to_wbart_text()
    cl = ff.cut_lists()
    ref = onp.forest_from_wbart(text, cl, offset=1.75)
    f = Forest.from_wbart(text, cl, offset=1.75)
    soa = f.export_soa()
    for k in ("tree_offset", "var", "cut", "left", "right"):
        assert np.array_equal(soa[k], getattr(ref, k)), k
    assert np.array_equal(soa["mu"], ref.mu)
    N = 1500
    X = rng.random((N, d))
    for j in range(d):                                       # a third of the points sit exactly on cut values
        X[j::3 * d, j] = rng.choice(cl[j], len(X[j::3 * d, j]))
    X[5, 0] = np.nan; X[6, :] = np.inf; X[7, :] = -np.inf; X[8, d - 1] = np.nextafter(cl[d - 1][3], np.inf)
    pts = f.points(X)
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(X, ref.cut_offsets, ref.cut_values, route=1))
    pts.close()
    want_leaf = onp.leaf_ordinals(ref, X)
    for table in ([False, True] if f.table_kernel_ok else [False]):
        assert np.array_equal(f.leaf_indices(X, use_table_kernel=table), want_leaf), table
    P = onp.predict(ref, X)
    w = rng.random(N)
    for walk in (False, True):
        out = f.eval(X, weights=w, want=("full_pred", "draw_mean", "point_mean", "point_var"), force_walk=walk)
        _close(out["full_pred"], P, 1e-13, 1.0)
        _close(out["draw_mean"], onp.row_means(P), 1e-13, 1.0)
        _close(out["point_mean"], P.mean(axis=0), 1e-12, 1.0)
        _close(out["point_var"], onp.col_vars(P) * w, 1e-10, float(np.max(onp.col_vars(P))))
    # the rule changes results on this input, and the same SoA forest built with route=1 agrees with the text path
    dbarts_rule = Forest.from_soa(ff)
    assert not np.array_equal(dbarts_rule.leaf_indices(X), want_leaf)
    routed = Forest.from_soa(FlatForest(ff.S, ff.m, ff.d, ff.tree_offset, ff.var, ff.cut, ff.left, ff.right, ff.mu,
                                        ff.cut_offsets, ff.cut_values, 1.25, 2.25, {}, 1))
    assert routed.route == 1 and np.array_equal(routed.leaf_indices(X), want_leaf)
    # exact integrals: identical under both rules (the rules differ on a null set); scaled units = no offset
    for measure in ("uniform", "exponential"):
        _close(f.exact_integrals(measure), dbarts_rule.exact_integrals(measure), 1e-15, 1.0)
        _close(f.exact_integrals(measure), onp.sample_integrals(ref, measure), 1e-12, 1.0)
    for h in (f, dbarts_rule, routed):
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("measure", ["uniform", "exponential"])
def test_wbart_device_point_generation(measure):
    rng = np.random.default_rng(3)
    d = 6
    ff = synth.prior_forest(3, 3, 4, rng.random((50, d)), None)
    f = Forest.from_wbart(ff.to_wbart_text(), ff.cut_lists())
    N, seed = 5000, 99
    import torch
    dX = torch.empty((d, N), dtype=torch.float64, device="cuda")
    pts = f.generate_points(measure, N, seed, d_x_out=dX.data_ptr())
    torch.cuda.synchronize()
    X = dX.cpu().numpy().T
    assert np.array_equal(X, onp.generate_points(measure, N, d, seed)) or np.allclose(X, onp.generate_points(measure, N, d, seed), rtol=1e-14, atol=0)
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(X, ff.cut_offsets, ff.cut_values, route=1))
    pts.close(); f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bad", ["", "1 1\n1\n1 0 0 0.5\n", "1 1 2\n2\n1 0 0 0\n2 0 0 1.0\n",
                                 "1 1 2\n3\n1 5 0 0\n2 0 0 1\n3 0 0 2\n", "1 1 2\n3\n1 0 9 0\n2 0 0 1\n3 0 0 2\n",
                                 "1 1 2\n1\n2 0 0 1.0\n", "1 1 2\n2\n1 0 0 1.0\n4 0 0 1.0\n", "1 1 3\n1\n1 0 0 1.0\n",
                                 "1 1 2\n1\n1 0 0 1.0\nextra\n"])
def test_wbart_rejects_malformed(bad):
    with pytest.raises(BartIntError):
        Forest.from_wbart(bad, KAT_CUTS, 0.0)


# ------------------------------------------------------------------------------------------------------------
# The C++ text parser runs before the library touches the device, so its verdicts are testable on a CPU-only box:
# malformed text -> BARTINT_ERR_PARSE (2) / _ARG (1); well-formed text -> accepted (a forest on a GPU box, else
# BARTINT_ERR_NODEVICE (6): there is no CPU path).
# ------------------------------------------------------------------------------------------------------------
def _from_wbart_status(text, cuts, offset=0.0):
    try:
        f = Forest.from_wbart(text, cuts, offset)
    except BartIntError as e:
        return e.status
    f.close()
    return 0


def test_wbart_cpp_parser_verdicts():
    from bartint_testlib import has_gpu
    accepted = 0 if has_gpu() else 6
    assert _from_wbart_status(KAT_TEXT, KAT_CUTS) == accepted
    assert _from_wbart_status(KAT_TEXT.replace("\n", "\r\n"), KAT_CUTS) == accepted          # CRLF line ends
    assert _from_wbart_status(KAT_TEXT.rstrip("\n"), KAT_CUTS) == accepted                    # no final newline
    assert _from_wbart_status("1 1 2\n1\n1 0 0 -1.5e-3\n", KAT_CUTS) == accepted              # exponent notation
    rng = np.random.default_rng(2)
    ff = synth.prior_forest(2, 5, 7, rng.random((30, 3)), None)
    assert _from_wbart_status(ff.to_wbart_text(), ff.cut_lists()) == accepted
    for bad in ["", "1 1 2\n2\n1 0 0 0\n2 0 0 1.0\n", "1 1 2\n3\n1 5 0 0\n2 0 0 1\n3 0 0 2\n",
                "1 1 2\n3\n1 0 9 0\n2 0 0 1\n3 0 0 2\n", "1 1 2\n1\n2 0 0 1.0\n", "1 1 2\n2\n1 0 0 1.0\n4 0 0 1.0\n",
                "1 1 2\n1\n1 0 0 1.0\nextra\n", "1 2 2\n1\n1 0 0 1.0\n", "1 1 2\n1\n1 0 0 abc\n", "1 1 2\n0\n",
                "1 1 2\n2\n1 0 0 1.0\n1 0 0 2.0\n", "-1 1 2\n1\n1 0 0 1.0\n"]:
        assert _from_wbart_status(bad, KAT_CUTS) == 2, repr(bad)
    assert _from_wbart_status("1 1 3\n1\n1 0 0 1.0\n", KAT_CUTS) == 1                         # p != number of cut lists
    assert _from_wbart_status("1 1\n1\n1 0 0 0.5\n", KAT_CUTS) == 1                            # header one token short -> p mismatch


# ---- the C++ exporter on the host (no GPU): bartint_debug_export_wbart against the oracle ----------------------------

@pytest.mark.parametrize("d", [1, 3, 10, 20])
def test_wbart_cpp_exporter_matches_oracle_on_host(d):
    import hosthooks
    rng = np.random.default_rng(300 + d)
    ff = synth.prior_forest(11 + d, 9, 14, rng.random((20 * d, d)), None)
    for text in (ff.to_wbart_text(), ff.to_wbart_text().replace("\n", "\r\n")):
        ref = onp.forest_from_wbart(text, ff.cut_lists(), offset=0.0)
        got = hosthooks.export_wbart(text, ff.cut_lists())
        assert (got["S"], got["m"], got["d"]) == (ref.S, ref.m, ref.d)
        for k in ("tree_offset", "var", "cut", "left", "right"):
            assert np.array_equal(got[k], getattr(ref, k)), k
        assert np.array_equal(got["mu"], ref.mu)


def test_wbart_cpp_exporter_kat_on_host():
    import hosthooks
    got = hosthooks.export_wbart(KAT_TEXT, KAT_CUTS)
    ref = onp.forest_from_wbart(KAT_TEXT, KAT_CUTS, offset=KAT_OFFSET)
    for k in ("tree_offset", "var", "cut", "left", "right", "mu"):
        assert np.array_equal(got[k], getattr(ref, k)), k


def test_mutated_wbart_text_never_crashes_the_parser():
    """Seeded byte-level mutations of a valid treedraws text: a parse / argument error or a structurally valid forest."""
    import hosthooks
    from bartint_b200 import BartIntError
    rng = np.random.default_rng(55)
    ff = synth.prior_forest(55, 3, 5, rng.random((30, 3)), None)
    text = ff.to_wbart_text()
    alphabet = list("0123456789 \n\n.-e+x\r")
    parsed = rejected = 0
    for it in range(400):
        chars = list(text)
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(chars)))
            op = int(rng.integers(4))
            if op == 0:
                del chars[pos]
            elif op == 1:
                chars.insert(pos, alphabet[int(rng.integers(len(alphabet)))])
            elif op == 2:
                chars[pos] = alphabet[int(rng.integers(len(alphabet)))]
            else:
                del chars[pos:pos + int(rng.integers(1, 40))]              # a truncated / spliced file
        try:
            got = hosthooks.export_wbart("".join(chars), ff.cut_lists())
        except BartIntError as e:
            assert e.status in (1, 2), e
            rejected += 1
            continue
        parsed += 1
        off = got["tree_offset"]
        assert len(off) == got["S"] * got["m"] + 1 and ((np.diff(off) % 2) == 1).all()
        internal = got["var"] >= 0
        assert (got["var"][internal] < ff.d).all()
    assert rejected > 100 and parsed > 10
