"""CPU tests of the tree-state exporter (forest_host.cu: parse_dbarts_forest -- the replacement of getTree,
src/BARTInt.R:92-133) through the host-only hook bartint_debug_export_dbarts098: the flattened forest must equal
the oracle's restatement node for node, for the prior's trees and for the shapes that stress the leaf-value search
(deep chains, leaves that share a value, leaves no training row reaches, tiny designs)."""
import numpy as np
import pytest

import hosthooks
from bartint_b200 import BartIntError, FlatForest, WireStrings, synth
from oracle import oracle_np as onp


def _export(ff, x_train, strings=None, fits=None):
    s2, f2 = ff.to_dbarts_state(x_train)
    strings = s2 if strings is None else strings
    fits = f2 if fits is None else fits
    nn, got = hosthooks.export_dbarts(WireStrings(strings), fits, x_train, ff.cut_offsets, ff.cut_values)
    want = onp.forest_from_dbarts(strings, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max)
    return nn, got, want


def _assert_same(got, want):
    assert np.array_equal(got["tree_offset"], want["tree_offset"])
    for k in ("var", "cut", "left", "right"):
        assert np.array_equal(got[k], want[k]), k
    leaf = want["var"] < 0
    assert np.array_equal(got["mu"][leaf], want["mu"][leaf], equal_nan=True)


@pytest.mark.parametrize("d,n,seed", [(1, 20, 0), (3, 40, 1), (10, 200, 2), (16, 7, 3), (20, 400, 4)])
def test_prior_forests_match_the_oracle(d, n, seed):
    rng = np.random.default_rng(seed)
    x_train = rng.random((n, d))
    ff = synth.prior_forest(seed, 6, 25, x_train, synth.genz_gaussian(x_train))
    nn, got, want = _export(ff, x_train)
    assert nn == ff.n_nodes
    _assert_same(got, want)
    # and the exporter inverts to_dbarts_state: same trees, same leaf values
    assert np.array_equal(got["var"], ff.var) and np.array_equal(got["cut"], ff.cut)
    assert np.array_equal(got["mu"][ff.var < 0], ff.mu[ff.var < 0])


@pytest.mark.parametrize("kind", ["stumps", "left_chain", "right_chain", "same_var", "balanced", "edge_cuts"])
def test_adversarial_shapes_including_unreached_leaves(kind):
    # 30 rows against chains of depth 9: many leaves receive no training row (-> NaN like the oracle) and most of
    # the others are found only by the fallback scan, not by the candidate rows
    ff = synth.adversarial_forest(kind, 3, 5, 4, numcut=50, seed=2, depth=9)
    rng = np.random.default_rng(9)
    x_train = rng.random((30, 4))
    _, got, want = _export(ff, x_train)
    _assert_same(got, want)


def test_leaves_sharing_one_value_and_duplicate_rows():
    cl = [np.linspace(0, 1, 12)[1:-1]] * 2
    trees = [(0, 4, 0.25, (1, 6, 0.25, 0.25)), (1, 2, (0, 7, -1.0, -1.0), 3.0)]
    ff = synth.forest_from_trees(trees * 2, 2, 2, cl)
    x_train = np.array([[0.1, 0.1], [0.1, 0.1], [0.9, 0.9], [0.9, 0.2], [0.45, 0.95], [0.2, 0.7]])
    _, got, want = _export(ff, x_train)
    _assert_same(got, want)
    assert not np.isnan(got["mu"][got["var"] < 0]).any()


def test_routing_ties_and_nan_rows():
    # x == cut goes left, NaN goes left (dbarts' integer cut map: code = #{cuts < x}, NaN -> 0)
    cl = [np.array([0.25, 0.5, 0.75])]
    ff = synth.forest_from_trees([(0, 1, 1.0, 2.0)], 1, 1, cl)
    for x, mu_left_seen in (([[0.5]], True), ([[np.nan]], True), ([[np.nextafter(0.5, 1)]], False)):
        x_train = np.array(x)
        fits = np.array([7.0]).reshape(1, 1, 1)
        _, got, want = _export(ff, x_train, fits=fits)
        _assert_same(got, want)
        assert (got["mu"][1] == 7.0) == mu_left_seen and (got["mu"][2] == 7.0) == (not mu_left_seen)


def test_no_training_rows():
    ff = synth.forest_from_trees([(0, 0, 1.0, 2.0), 0.5], 1, 2, [np.array([0.5])])
    nn, got, _ = _export(ff, np.zeros((0, 1)), fits=np.zeros((0, 2, 1)))
    assert nn == 4 and np.isnan(got["mu"][got["var"] < 0]).all()


@pytest.mark.parametrize("bad", ["", "0 1 .", "0 1 . . .", "0 x . .", "0 1 . . 3", "5 0 . .", "0 99 . .", "-1 0 . ."])
def test_malformed_strings_are_parse_errors(bad):
    strings = np.array([[bad]], dtype=object)
    with pytest.raises(BartIntError) as e:
        hosthooks.export_dbarts(WireStrings(strings), np.zeros((2, 1, 1)), np.zeros((2, 1)), [0, 3], [0.25, 0.5, 0.75])
    assert e.value.status == 2


def test_thread_count_does_not_change_the_result():
    from bartint_b200 import _lib
    rng = np.random.default_rng(5)
    x_train = rng.random((120, 6))
    ff = synth.prior_forest(5, 40, 30, x_train)
    lib = _lib.load()
    outs = []
    for t in (1, 3, 8):
        lib.bartint_set_host_threads(t)
        outs.append(_export(ff, x_train)[1])
    lib.bartint_set_host_threads(0)
    for o in outs[1:]:
        for k in ("tree_offset", "var", "cut", "left", "right"):
            assert np.array_equal(o[k], outs[0][k])
        assert np.array_equal(o["mu"], outs[0]["mu"], equal_nan=True)


def test_empty_and_degenerate_forests_on_the_host():
    from bartint_b200 import FlatForest
    z32 = np.zeros(0, np.int32)
    ff = FlatForest(0, 5, 2, np.zeros(1, np.int64), z32, z32, z32, z32, np.zeros(0), np.array([0, 1, 2], np.int32),
                    np.array([0.5, 0.5]), -0.5, 0.5)
    plan = hosthooks.plan_soa(ff)
    assert not plan["table_ok"] and plan["n_groups"] == 0                 # no trees: walk kernel territory, nothing planned
    nn, out = hosthooks.export_dbarts(WireStrings(np.empty((5, 0), dtype=object)), np.zeros((3, 5, 0)), np.zeros((3, 2)),
                                      [0, 1, 2], [0.5, 0.5])
    assert nn == 0 and list(out["tree_offset"]) == [0]
    stumps = synth.forest_from_trees([0.1, 0.2], 1, 2, [])                # d = 0: a forest of constants
    plan = hosthooks.plan_soa(stumps)
    assert plan["table_ok"] and not plan["gather"] and plan["n_groups"] == 1


def test_mutated_tree_strings_never_crash_the_parser():
    """Seeded byte-level mutations of valid dbarts tree strings (deletions, insertions, digit / dot / space swaps):
    the exporter either reports a parse error or returns a structurally valid forest."""
    rng = np.random.default_rng(77)
    x_train = rng.random((25, 3))
    ff = synth.prior_forest(77, 4, 6, x_train)
    strings, fits = ff.to_dbarts_state(x_train)
    alphabet = list("0123456789 ..  -x")
    parsed = rejected = 0
    for it in range(400):
        s2 = strings.copy()
        t, s = int(rng.integers(ff.m)), int(rng.integers(ff.S))
        chars = list(str(s2[t, s]))
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(chars) + 1))
            op = int(rng.integers(3))
            if op == 0 and chars:
                del chars[min(pos, len(chars) - 1)]
            elif op == 1:
                chars.insert(pos, alphabet[int(rng.integers(len(alphabet)))])
            elif chars:
                chars[min(pos, len(chars) - 1)] = alphabet[int(rng.integers(len(alphabet)))]
        s2[t, s] = "".join(chars)
        try:
            nn, got = hosthooks.export_dbarts(WireStrings(s2), fits, x_train, ff.cut_offsets, ff.cut_values)
        except BartIntError as e:
            assert e.status == 2, e
            rejected += 1
            continue
        parsed += 1
        off = got["tree_offset"]
        assert off[-1] == nn and ((np.diff(off) % 2) == 1).all()
        internal = got["var"] >= 0
        assert (got["var"][internal] < ff.d).all() and (got["left"][internal] > 0).all() and (got["right"][internal] > 0).all()
    assert rejected > 100 and parsed > 10


def test_draw_slices_export_to_the_matching_part_of_the_forest():
    """dist.sharded_design_step lets rank r export only draws [lo_r, hi_r): a WireStrings view plus the matching slab of
    savedTreeFits.  The pieces must be exactly the corresponding trees of the whole export."""
    rng = np.random.default_rng(31)
    x_train = rng.random((40, 5))
    ff = synth.prior_forest(31, 11, 7, x_train)
    strings, fits = ff.to_dbarts_state(x_train)
    fits = np.asfortranarray(np.asarray(fits))
    wire = WireStrings(strings)
    _, whole = hosthooks.export_dbarts(wire, fits, x_train, ff.cut_offsets, ff.cut_values)
    off = whole["tree_offset"]
    from bartint_b200.dist import shard_range
    covered = 0
    for rank in range(4):
        lo, hi = shard_range(ff.S, 4, rank)
        _, part = hosthooks.export_dbarts(wire.slice_draws(lo, hi), fits[:, :, lo:hi], x_train, ff.cut_offsets,
                                          ff.cut_values)
        a, b = off[lo * ff.m], off[hi * ff.m]
        assert np.array_equal(part["tree_offset"], off[lo * ff.m:hi * ff.m + 1] - a)
        for k in ("var", "cut", "left", "right"):
            assert np.array_equal(part[k], whole[k][a:b]), k
        assert np.array_equal(part["mu"], whole["mu"][a:b], equal_nan=True)
        covered += hi - lo
    assert covered == ff.S


@pytest.mark.parametrize("seed,S,m,d,n", [(1, 6, 5, 3, 40), (2, 33, 50, 10, 200), (3, 4, 7, 1, 21), (4, 3, 9, 20, 60)])
def test_one_pass_walk_exporter_equals_two_pass(seed, S, m, d, n):
    """The fused design step parses dbarts strings straight into the walk form (parse_dbarts_walk); it must produce
    exactly what the exporter + validation + walk-form packer produce: node words, leaf offsets, leaf values (getTree's
    rule, BARTInt.R:92-133) and the statistics the bit-sliced kernel is sized from."""
    rng = np.random.default_rng(seed)
    x_train = rng.random((n, d))
    ff = synth.prior_forest(seed, S, m, x_train, synth.genz_gaussian(x_train))
    strings, fits = ff.to_dbarts_state(x_train)
    wire = WireStrings(strings)
    nn, two = hosthooks.export_dbarts(wire, fits, x_train, ff.cut_offsets, ff.cut_values)
    flat = FlatForest(S, m, d, two["tree_offset"], two["var"], two["cut"], two["left"], two["right"], two["mu"],
                      ff.cut_offsets, ff.cut_values, ff.y_min, ff.y_max)
    plan = hosthooks.plan_soa(flat)
    one = hosthooks.export_dbarts_walk(wire, fits, x_train, ff.cut_offsets, ff.cut_values, nn)
    assert int(one["stats"][0]) == nn and int(one["stats"][1]) == (nn + S * m) // 2
    assert np.array_equal(one["walk_nodes"], plan["walk_nodes"][:2 * nn])
    assert np.array_equal(one["tree_off"], two["tree_offset"].astype(np.int32))
    assert np.array_equal(one["leaf_off"], plan["leaf_off"])
    assert np.array_equal(one["leaf_mu"], plan["leaf_mu"][:len(one["leaf_mu"])], equal_nan=True)
    internal = two["var"] >= 0
    per_draw = [int(internal[two["tree_offset"][s * m]:two["tree_offset"][(s + 1) * m]].sum()) for s in range(S)]
    assert int(one["stats"][4]) == max(per_draw)
    # malformed strings fail the same way
    bad = strings.copy()
    bad[0, 0] = "0 1 . "
    with pytest.raises(BartIntError) as e:
        hosthooks.export_dbarts_walk(WireStrings(bad), fits, x_train, ff.cut_offsets, ff.cut_values, nn)
    assert e.value.status == 2
