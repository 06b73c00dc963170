"""GPU parity of the bit-sliced sums-only kernel (kernels_bitslice.cu) against the oracle: rowMeans(predict(model, X))
of src/survey_design/bartMean.R:74-82 / the Monte Carlo integral, for every tile shape (8 / 4 / 2 chunks of 128 points),
tree shapes the prior rarely produces (deep chains, balanced trees, nested cuts of one variable, stumps), ragged
point counts, draw sub-ranges that start and end inside a slice of 32 draws, both routing rules, and the three code
sources (packed records, feature-major rows, points generated on the device)."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp
from bartint_b200 import Forest, synth

pytestmark = pytest.mark.gpu


def _close(got, want, tol, scale=1.0):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape
    err = float(np.max(np.abs(got - want))) if got.size else 0.0
    assert err <= tol * scale, err / scale


def _check(ff, X, draws=None, tol=1e-13):
    """draw means of the bit-sliced kernel == oracle (long double row means) == the per-point kernels"""
    f = Forest.from_soa(ff)
    assert f.bitslice_ok
    scale = max(ff.y_max - ff.y_min, 1e-300)
    ref = oc.predict_reduce(ff, X, nthreads=oc.max_threads(), want=("draw_mean",))["draw_mean"]
    if draws is not None:
        ref = ref[draws[0]:draws[1]]
    got = f.eval(X, want=("draw_mean",), draws=draws)["draw_mean"]
    assert f.last_eval_profile(False)["kernel"] == "bitslice"
    _close(got, ref, tol, scale)
    other = f.eval(X, want=("draw_mean",), draws=draws, no_bitslice=True)["draw_mean"]
    assert f.last_eval_profile(False)["kernel"] != "bitslice"
    _close(got, other, 1e-12, scale)
    f.close()


@pytest.mark.parametrize("d,numcut,S,m,N,chunks", [
    (1, 100, 70, 50, 1500, 8),        # cfg 1 shape; S is not a multiple of 32
    (3, 100, 33, 10, 31, 8),          # fewer points than one 32-point word
    (10, 100, 64, 50, 20011, 8),      # cfg 2 shape, ragged last tile
    (20, 100, 40, 200, 3001, 4),      # cfg 3 shape: 2000 predicate rows -> 512-point tiles
    (40, 120, 35, 20, 2500, 2),       # 4800 rows -> 256-point tiles
    (12, 255, 20, 30, 4097, 4),       # the largest cut count a uint8 code allows
])
def test_prior_forests(d, numcut, S, m, N, chunks):
    rng = np.random.default_rng(1000 + d)
    x_train = rng.random((max(5 * d, 30), d))
    ff = synth.prior_forest(17 + d, S, m, x_train, synth.genz_gaussian(x_train), numcut=numcut)
    X = rng.random((N, d))
    cl = ff.cut_lists()
    for j in range(d):                                   # points exactly on cut values and just above them
        X[j::97, j] = rng.choice(cl[j], len(X[j::97, j]))
        X[j + 40::97, j] = np.nextafter(rng.choice(cl[j], len(X[j + 40::97, j])), np.inf)
    if N > 8:
        X[3, 0] = np.nan; X[4, :] = np.inf; X[5, :] = -np.inf
    f = Forest.from_soa(ff)
    assert f.bitslice_ok and f.bitslice_tile == 128 * chunks, f.bitslice_info
    f.close()
    _check(ff, X)


@pytest.mark.parametrize("kind,depth", [("left_chain", 12), ("right_chain", 9), ("balanced", 3), ("same_var", 5),
                                        ("edge_cuts", 2), ("left_chain", 30)])
def test_adversarial_shapes(kind, depth):
    ff = synth.adversarial_forest(kind, 37, 7, 4, seed=depth, depth=depth)
    X = np.random.default_rng(depth).random((5003, 4))
    _check(ff, X)


def test_stumps_and_mixed():
    """A forest of stumps has nothing to slice (the per-point kernels take it); stumps inside ordinary draws do."""
    ff = synth.adversarial_forest("stumps", 5, 4, 2)
    f = Forest.from_soa(ff)
    assert not f.bitslice_ok
    X = np.random.default_rng(0).random((100, 2))
    _close(f.eval(X, want=("draw_mean",))["draw_mean"], onp.row_means(onp.predict(ff, X)), 1e-14)
    f.close()
    rng = np.random.default_rng(8)
    x_train = rng.random((30, 3))
    ff = synth.prior_forest(3, 50, 12, x_train, base=0.3)          # base 0.3: most trees are stumps
    _check(ff, rng.random((7000, 3)))


def test_draw_subranges_and_point_sources():
    rng = np.random.default_rng(4)
    x_train = rng.random((60, 5))
    ff = synth.prior_forest(9, 100, 20, x_train, synth.genz_gaussian(x_train))
    X = rng.random((9001, 5))
    for draws in [(0, 100), (0, 1), (31, 33), (40, 97), (64, 96), (99, 100)]:
        _check(ff, X, draws=draws)
    # feature-major code rows (table forests keep no packed records)
    import os
    os.environ["BARTINT_KERNEL"] = "table"
    try:
        _check(ff, X)
    finally:
        del os.environ["BARTINT_KERNEL"]
    # points generated on the device: same per-draw means as the per-point kernels on the same point set
    f = Forest.from_soa(ff)
    pts = f.generate_points("uniform", 50_000, seed=5)
    a = f.eval(pts, want=("draw_mean",))["draw_mean"]
    b = f.eval(pts, want=("draw_mean",), no_bitslice=True)["draw_mean"]
    _close(a, b, 1e-13, ff.y_max - ff.y_min)
    pts.close(); f.close()


@pytest.mark.parametrize("d,numcut,S,m,N,chunks", [
    (4, 100, 1100, 6, 2500, 8),       # two permutation blocks of 1024 draws
    (3, 60, 2500, 3, 1300, 8),        # three blocks, the last one partial
    (12, 255, 300, 5, 1500, 4),       # masks leave room for 256 accumulators only: blocks of 256 draws
])
def test_permutation_blocks(d, numcut, S, m, N, chunks):
    """Each class sorts the draws inside blocks of 1024 (or 256) draws; a CTA owns one block.  Draw ranges that start or
    end inside a block, or cover a block none of whose slices is wanted, must give the same per-draw means."""
    rng = np.random.default_rng(77 + S)
    x_train = rng.random((max(5 * d, 30), d))
    ff = synth.prior_forest(5 + d, S, m, x_train, synth.genz_gaussian(x_train), numcut=numcut)
    X = rng.random((N, d))
    f = Forest.from_soa(ff)
    assert f.bitslice_ok and f.bitslice_tile == 128 * chunks, f.bitslice_info
    f.close()
    _check(ff, X)
    for draws in [(0, 256), (255, 257), (S - 40, S), (200, S - 1), (1023, min(1025, S))]:
        if draws[1] <= S and draws[0] < draws[1]:
            _check(ff, X, draws=draws)


def test_leaf_value_scales():
    """The fixed-point scale follows max |mu|: tiny, huge and mixed-magnitude leaf values keep ~1e-16 relative accuracy."""
    rng = np.random.default_rng(2)
    x_train = rng.random((40, 3))
    base = synth.prior_forest(21, 40, 15, x_train)
    X = rng.random((6007, 3))
    unit = oc.predict_reduce(base, X, nthreads=4, want=("draw_mean",))["draw_mean"]   # y range [-0.5, 0.5]: scaled units
    for factor in (2.0 ** -660, 1e-9, 1.0, 3e7, 2.0 ** 500):
        ff = synth.FlatForest(base.S, base.m, base.d, base.tree_offset, base.var, base.cut, base.left, base.right,
                              base.mu * factor, base.cut_offsets, base.cut_values, -0.5, 0.5)
        f = Forest.from_soa(ff)
        assert f.bitslice_ok
        got = f.eval(X, want=("draw_mean",), scaled=True)["draw_mean"]
        assert f.last_eval_profile(False)["kernel"] == "bitslice"
        _close(got, unit * factor, 1e-13, float(np.max(np.abs(base.mu))) * factor * base.m)
        f.close()


def test_wbart_routing_rule():
    """BART::wbart routing (left iff x < cut) only changes the binning; the bit-sliced kernel works on codes."""
    rng = np.random.default_rng(6)
    x_train = rng.random((50, 4))
    ff = synth.prior_forest(31, 48, 10, x_train)
    ff.route = 1
    X = rng.random((4000, 4))
    cl = ff.cut_lists()
    X[::7, 1] = rng.choice(cl[1], len(X[::7, 1]))
    _check(ff, X)
