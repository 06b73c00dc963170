"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): leaf indices and cut codes BIT-EXACT; integrals, means and variances
within 1e-6 relative (we assert much tighter: the arithmetic is fp64 end to end, differences come only
from summation order).
"""
import math

import numpy as np
import pytest

from oracle import oracle_np as onp
from bartint_b200 import Forest, synth, api
from bartint_testlib import apply_kernel_kind, kat_forest, kernel_kind_params

pytestmark = pytest.mark.gpu

@pytest.fixture(params=kernel_kind_params(), autouse=True)
def kernel_kind(request, monkeypatch):
    """Every parity test runs against both fast kernels: the register-gather kernel (default for d <= 16) and the
    shared-memory table kernel (forced with BARTINT_KERNEL=table; the default for d > 16) -- and against the gather
    kernel in counting mode (BARTINT_COUNT_SPLITS=1: single-split trees' per-draw sums from code histograms)."""
    return apply_kernel_kind(request.param, monkeypatch)


RTOL = 1e-6            # the north-star tolerance
TIGHT = 1e-11          # what fp64 accumulation should actually give


def _assert_close(got, want, tol=TIGHT, scale=None, msg=""):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape, msg
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), msg
    s = np.maximum(np.abs(want[~nan]), 1e-300) if scale is None else scale
    err = np.abs(got[~nan] - want[~nan]) / s
    assert err.size == 0 or err.max() <= tol, f"{msg}: max rel err {err.max():.3e}"


def test_kats_cuda(kats):
    for k in kats:
        ff = kat_forest(k)
        X = np.array(k["X"], dtype=float)
        f = Forest.from_soa(ff)
        for table in ([False, True] if f.table_kernel_ok else [False]):
            assert np.array_equal(f.leaf_indices(X, use_table_kernel=table), np.array(k["leaf"])), (k["name"], table)
        for walk in (False, True):
            out = f.eval(X, want=("full_pred", "draw_mean", "point_var"), force_walk=walk)
            _assert_close(out["full_pred"], k["predict"], 1e-15, msg=k["name"])
            if "col_vars" in k:
                _assert_close(out["point_var"], k["col_vars"], 1e-13, msg=k["name"])
                _assert_close(out["draw_mean"], k["row_means"], 1e-14, msg=k["name"])
        for measure, want in k["integrals"].items():
            _assert_close(f.exact_integrals(measure), want, 1e-14, scale=1.0, msg=f"{k['name']} {measure}")
        f.close()


def test_codes_bit_exact(small_case):
    ff, _, _, X = small_case
    Xe = X.copy()
    cl = ff.cut_lists()
    Xe[0, 0] = cl[0][10]; Xe[1, 0] = np.nextafter(cl[0][10], np.inf); Xe[2, 1] = -5.0; Xe[3, 1] = 9.0
    Xe[4, 2] = np.nan; Xe[5, 3] = cl[3][0]; Xe[6, 3] = cl[3][-1]; Xe[7, 3] = np.nextafter(cl[3][-1], np.inf)
    Xe[8, :] = np.inf; Xe[9, :] = -np.inf
    f = Forest.from_soa(ff)
    pts = f.points(Xe)
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(Xe, ff.cut_offsets, ff.cut_values))
    pts.close(); f.close()


@pytest.mark.parametrize("table", [False, True])
def test_leaf_indices_bit_exact(small_case, table):
    ff, _, _, X = small_case
    f = Forest.from_soa(ff)
    if table and not f.table_kernel_ok:
        pytest.skip("table kernel not applicable")
    got = f.leaf_indices(X, use_table_kernel=table)
    want = onp.leaf_ordinals(ff, X)
    assert got.shape == want.shape and np.array_equal(got, want)
    f.close()


@pytest.mark.parametrize("walk", [False, True])
def test_predict_and_reductions(small_case, walk):
    ff, _, _, X = small_case
    f = Forest.from_soa(ff)
    w = np.random.default_rng(1).random(X.shape[0])
    out = f.eval(X, weights=w, want=("full_pred", "draw_mean", "point_mean", "point_var"), force_walk=walk)
    P = onp.predict(ff, X)
    _assert_close(out["full_pred"], P, 1e-13, scale=ff.y_max - ff.y_min, msg="predict")
    _assert_close(out["draw_mean"], onp.row_means(P), 1e-13, msg="rowMeans")
    _assert_close(out["point_mean"], P.mean(axis=0), 1e-12, msg="colMeans")
    _assert_close(out["point_var"], onp.col_vars(P) * w, 1e-10, msg="colVars*weights")
    sc = f.eval(X, want=("full_pred",), scaled=True, force_walk=walk)
    _assert_close(sc["full_pred"], onp.tree_sums(ff, X), 1e-13, scale=1.0, msg="scaled units")
    f.close()


@pytest.mark.parametrize("bits", [4, 5, 6, 7, 8])
def test_table_bits_variants(small_case, bits, monkeypatch, kernel_kind):
    ff, _, _, X = small_case
    monkeypatch.setenv("BARTINT_TABLE_BITS", str(bits))
    f = Forest.from_soa(ff)
    assert f.table_bits == (min(bits, 5) if kernel_kind == "gather" else bits)
    if not f.table_kernel_ok:
        pytest.skip("forest does not fit")
    out = f.eval(X, want=("draw_mean", "point_var"))
    P = onp.predict(ff, X)
    _assert_close(out["draw_mean"], onp.row_means(P), 1e-13)
    _assert_close(out["point_var"], onp.col_vars(P), 1e-10)
    assert np.array_equal(f.leaf_indices(X, use_table_kernel=True), onp.leaf_ordinals(ff, X))
    assert f.last_eval_profile(True)["table_kernel"]
    f.close()


@pytest.mark.parametrize("kind", ["stumps", "left_chain", "right_chain", "same_var", "balanced", "edge_cuts"])
def test_adversarial_forests(kind):
    ff = synth.adversarial_forest(kind, 5, 7, 3, numcut=100, seed=2, depth=11)
    rng = np.random.default_rng(6)
    X = rng.random((700, 3))
    X[:50] = np.round(X[:50] * 101) / 101          # many exact ties with the cut grid
    f = Forest.from_soa(ff)
    assert np.array_equal(f.leaf_indices(X), onp.leaf_ordinals(ff, X))
    if f.table_kernel_ok:
        assert np.array_equal(f.leaf_indices(X, use_table_kernel=True), onp.leaf_ordinals(ff, X))
    P = onp.predict(ff, X)
    for walk in (False, True):
        out = f.eval(X, want=("draw_mean", "point_var", "full_pred"), force_walk=walk)
        _assert_close(out["full_pred"], P, 1e-13, scale=1.0, msg=kind)      # y range is 1; values cross zero
        _assert_close(out["draw_mean"], onp.row_means(P), 1e-13, scale=1.0, msg=kind)
        _assert_close(out["point_var"], onp.col_vars(P), 1e-9, scale=np.maximum(onp.col_vars(P), 1e-12), msg=kind)
    _assert_close(f.exact_integrals("uniform"), onp.sample_integrals(ff, "uniform"), 1e-13, scale=1.0, msg=kind)
    f.close()


@pytest.mark.parametrize("measure", ["uniform", "gaussian", "exponential", "gaussian_correct"])
def test_exact_integrals(small_case, measure):
    ff, *_ = small_case
    f = Forest.from_soa(ff)
    want = onp.sample_integrals(ff, measure)
    _assert_close(f.exact_integrals(measure), want, 1e-12, scale=np.maximum(np.abs(want), 1e-3), msg=measure)
    # reference quirk B1: only the first min(m, S) draws
    n = min(ff.m, ff.S)
    _assert_close(f.exact_integrals(measure, n_draws_used=n), want[:n], 1e-12, scale=np.maximum(np.abs(want[:n]), 1e-3))
    res, mean, sd = f.finalize_integrals(want)
    r0, m0, s0 = onp.finalize_integrals(want, ff.y_min, ff.y_max)
    _assert_close(res, r0, 1e-14)
    assert abs(mean - m0) <= 1e-13 * max(1, abs(m0)) and abs(sd - s0) <= 1e-12 * max(1e-3, abs(s0))
    f.close()


def test_exporter_matches_oracle_gettree(small_case):
    ff, x_train, *_ = small_case
    strings, fits = ff.to_dbarts_state(x_train)
    f = Forest.from_dbarts(strings, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max)
    got = f.export_soa()
    want = onp.forest_from_dbarts(strings, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max)
    for key in ("tree_offset", "var", "cut", "left", "right", "mu"):
        assert np.array_equal(got[key], np.asarray(want[key])), key
    f.close()


def test_edge_shapes():
    rng = np.random.default_rng(2)
    x_train = rng.random((15, 1))
    # d = 1, S = 1 (variance NaN like R), m = 1, C = 1
    ff = synth.prior_forest(3, 1, 1, x_train)
    f = Forest.from_soa(ff)
    X = rng.random((1, 1))
    out = f.eval(X, want=("draw_mean", "point_mean", "point_var", "full_pred"))
    P = onp.predict(ff, X)
    _assert_close(out["full_pred"], P, 1e-14)
    assert math.isnan(out["point_var"][0])
    # empty point set: rowMeans of a 0-column matrix is NaN
    out = f.eval(np.zeros((0, 1)), want=("draw_mean",))
    assert np.isnan(out["draw_mean"]).all()
    f.close()
    # ragged: N not a multiple of anything, draw sub-range
    ff = synth.prior_forest(4, 9, 5, rng.random((20, 2)))
    f = Forest.from_soa(ff)
    X = rng.random((2049 * 2 + 1, 2))
    P = onp.predict(ff, X)
    for walk in (False, True):
        out = f.eval(X, want=("draw_mean", "point_var"), draws=(2, 7), force_walk=walk)
        _assert_close(out["draw_mean"], onp.row_means(P[2:7]), 1e-13)
        _assert_close(out["point_var"], onp.col_vars(P[2:7]), 1e-10, scale=np.maximum(onp.col_vars(P[2:7]), 1e-12))
    f.close()


def test_deep_tree_inside_table_kernel():
    """A tree with more than NB internal nodes becomes a WALK record inside the table kernel."""
    rng = np.random.default_rng(3)
    cut_lists = [np.linspace(0, 1, 102)[1:-1]] * 2
    deep = synth.adversarial_forest("left_chain", 1, 1, 2, seed=5, depth=14)
    a = synth.prior_forest(6, 4, 6, rng.random((25, 2)), cut_lists=cut_lists)
    # splice the deep tree in as tree 2 of draw 1
    trees = []
    for k in range(a.S * a.m):
        src, kk = (deep, 0) if k == 1 * a.m + 2 else (a, k)
        lo, hi = src.tree_offset[kk], src.tree_offset[kk + 1]
        trees.append((src.var[lo:hi], src.cut[lo:hi], src.left[lo:hi], src.right[lo:hi], src.mu[lo:hi]))
    off = np.concatenate([[0], np.cumsum([len(t[0]) for t in trees])])
    ff = synth.FlatForest(a.S, a.m, 2, off, *(np.concatenate([t[i] for t in trees]) for i in range(5)),
                          a.cut_offsets, a.cut_values, 0.0, 2.0)
    f = Forest.from_soa(ff)
    assert f.table_kernel_ok and f.max_internal > f.table_bits
    X = rng.random((5000, 2))
    P = onp.predict(ff, X)
    out = f.eval(X, want=("draw_mean", "point_var", "full_pred"))
    assert f.last_eval_profile(True)["table_kernel"]
    _assert_close(out["full_pred"], P, 1e-13, scale=2.0)
    _assert_close(out["draw_mean"], onp.row_means(P), 1e-13, scale=2.0)
    assert np.array_equal(f.leaf_indices(X, use_table_kernel=True), onp.leaf_ordinals(ff, X))
    f.close()


def test_api_drop_in_semantics(small_case):
    """sampleIntegrals / predict / colVars*weights / rowMeans through the R-interface mirror."""
    ff, x_train, y_train, X = small_case
    model = api.model_from_flat(ff, x_train, y_train)
    got = api.sampleIntegrals(model, ff.d, "uniform")                 # bug-compatible: first min(m, S) draws
    want = onp.sample_integrals(ff, "uniform", n_draws_used=min(ff.m, ff.S))
    _assert_close(got, want, 1e-12, scale=np.maximum(np.abs(want), 1e-3))
    Xc = X[:100]
    P = onp.predict(ff, Xc)
    _assert_close(api.predict(model, Xc), P, 1e-13)
    _assert_close(api.bartint_candidate_var(model, Xc, 1.0), onp.col_vars(P), 1e-10)
    _assert_close(api.bartint_draw_means(model, Xc), onp.row_means(P), 1e-13)
    model.close()


def test_sequential_design_loop_runs():
    rng = np.random.default_rng(0)
    trainX = rng.random((20, 1))
    out = api.mainBARTInt(1, 2, lambda x: np.exp(-np.abs(x - 0.5).sum(axis=1)), trainX,
                          np.exp(-np.abs(trainX - 0.5).sum(axis=1)), sequential=True, measure="uniform",
                          bart=lambda X, y, **kw: api.synthetic_bart(X, y, **{**kw, "ndpost": kw["ndpost"] // 50}))
    assert out["meanValueBART"].shape == (2,) and np.isfinite(out["meanValueBART"]).all()
    assert out["trainData"].shape == (21, 2)
    out["model"].close()


def test_full_size_properties_cfg2_shape():
    """BASELINE cfg-2-shaped run at reduced S (size-independent properties): linearity of the per-draw
    sum in the point set, MC mean within 5 sigma of the exact integral, table == walk kernel."""
    cfg = synth.CONFIGS[2]
    x_train, y_train = synth.make_design(2)
    ff = synth.prior_forest(2002, 64, cfg["m"], x_train, y_train)
    f = Forest.from_soa(ff)
    X = synth.make_points(2, 200_000)
    a = f.eval(X, want=("draw_mean",), scaled=True)["draw_mean"]
    b = f.eval(X, want=("draw_mean",), scaled=True, force_walk=True)["draw_mean"]
    _assert_close(a, b, 1e-12, scale=np.maximum(np.abs(b), 1e-3), msg="table vs walk")
    h1 = f.eval(X[:120_001], want=("draw_mean",), scaled=True)["draw_mean"]
    h2 = f.eval(X[120_001:], want=("draw_mean",), scaled=True)["draw_mean"]
    _assert_close((h1 * 120_001 + h2 * (200_000 - 120_001)) / 200_000, a, 1e-12, scale=np.maximum(np.abs(a), 1e-3),
                  msg="additivity over point shards")
    exact = f.exact_integrals("uniform")
    # the cut grid spans the training range, a subset of [0,1]^d, so MC over U[0,1]^d estimates the same integral
    sd = np.sqrt(f.eval(X[:20000], want=("full_pred",), scaled=True)["full_pred"].var(axis=1, ddof=1) / 200_000)
    assert (np.abs(a - exact) <= 6 * sd + 1e-9).all()
    f.close()


def test_dist_wrappers_single_rank(small_case):
    """The production multi-GPU wrappers (device tensors, async ABI) at world size 1 == the host API."""
    import torch
    from bartint_b200 import dist as bdist
    ff, _, _, X = small_case
    f = Forest.from_soa(ff)
    dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()          # [d, N] == N x d column-major
    pts = f.points_from_device(dX.data_ptr(), X.shape[0])
    dm, mv = bdist.point_sharded_draw_means(f, pts)
    w = torch.rand(X.shape[0], dtype=torch.float64, device="cuda")
    var = bdist.draw_sharded_candidate_var(f, pts, w)
    torch.cuda.synchronize()
    P = onp.predict(ff, X)
    rm = onp.row_means(P)
    _assert_close(dm.cpu().numpy(), rm, 1e-13)
    _assert_close(mv.cpu().numpy(), [rm.mean(), rm.var(ddof=1)], 1e-11)
    _assert_close(var.cpu().numpy(), onp.col_vars(P) * w.cpu().numpy(), 1e-10)
    # async exact integral + fused finalisation
    d_int = torch.empty(ff.S, dtype=torch.float64, device="cuda")
    d_res = torch.empty(ff.S, dtype=torch.float64, device="cuda")
    d_ms = torch.empty(2, dtype=torch.float64, device="cuda")
    f.exact_integrals_device(d_int.data_ptr(), "uniform", d_rescaled=d_res.data_ptr(), d_mean_sd=d_ms.data_ptr())
    torch.cuda.synchronize()
    want = onp.sample_integrals(ff, "uniform")
    r0, m0, s0 = onp.finalize_integrals(want, ff.y_min, ff.y_max)
    _assert_close(d_int.cpu().numpy(), want, 1e-12, scale=np.maximum(np.abs(want), 1e-3))
    _assert_close(d_res.cpu().numpy(), r0, 1e-13)
    _assert_close(d_ms.cpu().numpy(), [m0, s0], 1e-11)
    pts.close(); f.close()


def test_codes_irregular_cut_grid():
    """K0's affine guess must never change the answer: quantile-like, clustered and duplicated cut points."""
    rng = np.random.default_rng(9)
    cuts = [np.sort(rng.beta(0.3, 0.3, 100)), np.sort(np.concatenate([rng.normal(0.2, 1e-3, 60), rng.random(40)])),
            np.sort(np.repeat(rng.random(25), 4)), np.array([0.5])]
    ff = synth.forest_from_trees([(j, 0, 0.1, -0.1) for j in range(4)], 1, 4, cuts)
    X = rng.random((50001, 4))
    X[:100, 0] = cuts[0]; X[100:200, 1] = cuts[1]; X[200:300, 2] = cuts[2]          # exact ties with every cut
    X[300:400, 0] = np.nextafter(cuts[0], np.inf); X[400:500, 1] = np.nextafter(cuts[1], -np.inf)
    X[500] = np.nan; X[501] = np.inf; X[502] = -np.inf; X[503] = 1e300; X[504] = -1e300
    f = Forest.from_soa(ff)
    pts = f.points(X)
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(X, ff.cut_offsets, ff.cut_values))
    pts.close(); f.close()


@pytest.mark.parametrize("measure", ["uniform", "gaussian", "exponential"])
def test_device_point_generation(small_case, measure):
    """Philox points drawn on the device: values match the oracle's restatement of the generator, codes are the
    exact cut map of those values, shards reproduce the unsharded stream, and the MC estimate built on them agrees
    with the exact integral of the same measure."""
    import torch
    ff, *_ = small_case
    f = Forest.from_soa(ff)
    N, seed = 20001, 0x1234_5678_9ABC_DEF0
    dX = torch.empty((ff.d, N), dtype=torch.float64, device="cuda")
    pts = f.generate_points(measure, N, seed, d_x_out=dX.data_ptr())
    torch.cuda.synchronize()
    X = dX.cpu().numpy().T
    want = onp.generate_points(measure, N, ff.d, seed)
    if measure == "uniform":
        assert np.array_equal(X, want)                                  # bit-exact
    else:
        np.testing.assert_allclose(X, want, rtol=1e-13, atol=1e-15)     # libm vs CUDA math: last-ulp differences
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(X, ff.cut_offsets, ff.cut_values))
    # two shards == one stream
    a = f.generate_points(measure, 12000, seed); b = f.generate_points(measure, N - 12000, seed, first_index=12000)
    assert np.array_equal(np.vstack([a.codes(), b.codes()]), pts.codes())
    # MC on generated points vs the exact integral of the same measure
    exact_name = {"uniform": "uniform", "gaussian": "gaussian_correct", "exponential": "exponential"}[measure]
    exact = f.exact_integrals(exact_name)
    F = onp.tree_sums(ff, X)
    mc = f.eval(pts, want=("draw_mean",), scaled=True)["draw_mean"]
    _assert_close(mc, F.mean(axis=1), 1e-12, scale=np.maximum(np.abs(F.mean(axis=1)), 1e-3))
    se = F.std(axis=1, ddof=1) / math.sqrt(N)
    assert (np.abs(mc - exact) <= 5 * se + 1e-12).all()
    for p in (pts, a, b):
        p.close()
    f.close()


@pytest.mark.parametrize("d", [1, 5, 8, 9, 12, 13, 16, 17, 24])
def test_dimension_sweep_both_kernels(d, kernel_kind):
    """The gather kernel's register layouts (2, 3, 4 code words per point) and the table kernel beyond 16 variables."""
    rng = np.random.default_rng(100 + d)
    x_train = rng.random((50, d))
    ff = synth.prior_forest(d, 7, 11, x_train, rng.normal(size=50))
    X = rng.random((4500, d))
    f = Forest.from_soa(ff)
    P = onp.predict(ff, X)
    out = f.eval(X, want=("draw_mean", "point_var", "full_pred"))
    assert f.last_eval_profile(True)["table_kernel"]
    span = ff.y_max - ff.y_min                 # the response is N(0,1): predictions cross zero, so scale by the range
    _assert_close(out["full_pred"], P, 1e-13, scale=span)
    _assert_close(out["draw_mean"], onp.row_means(P), 1e-13, scale=span)
    _assert_close(out["point_var"], onp.col_vars(P), 1e-9, scale=np.maximum(onp.col_vars(P), 1e-12))
    assert np.array_equal(f.leaf_indices(X[:600], use_table_kernel=True), onp.leaf_ordinals(ff, X[:600]))
    f.close()


def test_argmax_ties_device():
    import ctypes as C
    import torch
    from bartint_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for case in range(6):
        v = rng.random(1000 + case)
        if case == 1:
            v[[5, 700, 333]] = 2.0                    # ties
        if case == 2:
            v[17] = np.nan
        if case == 3:
            v[:] = 0.25                               # everything ties
        if case == 4:
            v = -rng.random(7) - 5.0                  # all negative, short
        t = torch.from_numpy(v).cuda()
        out = torch.zeros(2, dtype=torch.int64, device="cuda")
        assert lib.bartint_argmax_ties_device(C.c_void_p(t.data_ptr()), v.size, C.c_void_p(out.data_ptr()), None) == 0
        idx, ties = out.cpu().tolist()
        if np.isnan(v).any():
            assert (idx, ties) == (-1, 0)
        else:
            which = np.nonzero(v == v.max())[0]       # R: which(var == max(var))
            assert idx == which[0] and ties == len(which)


def test_random_shapes_property(kernel_kind):
    """Seeded sweep over ragged shapes: S, m, d, N, draw sub-ranges, weights -- CUDA == oracle every time."""
    rng = np.random.default_rng(2024)
    for trial in range(12):
        d = int(rng.integers(1, 21)); S = int(rng.integers(1, 9)); m = int(rng.integers(1, 13))
        N = int(rng.choice([1, 3, 31, 257, 2047, 2048, 2049, 4097, 6000]))
        x_train = rng.random((int(rng.integers(2, 60)), d))
        ff = synth.prior_forest(int(rng.integers(1 << 30)), S, m, x_train, rng.normal(size=len(x_train)),
                                numcut=int(rng.choice([1, 2, 7, 100, 127])))
        X = rng.random((N, d)) * 1.4 - 0.2
        f = Forest.from_soa(ff)
        lo = int(rng.integers(0, S)); hi = int(rng.integers(lo + 1, S + 1))
        w = rng.random(N)
        out = f.eval(X, weights=w, want=("draw_mean", "point_mean", "point_var"), draws=(lo, hi))
        P = onp.predict(ff, X)[lo:hi]
        span = max(ff.y_max - ff.y_min, 1e-300)
        _assert_close(out["draw_mean"], onp.row_means(P), 1e-12, scale=span, msg=f"trial {trial}")
        _assert_close(out["point_mean"], P.mean(axis=0), 1e-12, scale=span, msg=f"trial {trial}")
        if hi - lo > 1:
            _assert_close(out["point_var"], onp.col_vars(P) * w, 1e-9,
                          scale=np.maximum(onp.col_vars(P) * w, 1e-12 * span ** 2), msg=f"trial {trial}")
        else:
            assert np.isnan(out["point_var"]).all()
        assert np.array_equal(f.leaf_indices(X, draws=(lo, hi)), onp.leaf_ordinals(ff, X)[lo * m:hi * m])
        if f.table_kernel_ok:
            assert np.array_equal(f.leaf_indices(X, draws=(lo, hi), use_table_kernel=True),
                                  onp.leaf_ordinals(ff, X)[lo * m:hi * m])
        f.close()


def test_pipelined_host_eval_matches_point_handle_path():
    """bartint_eval (host matrix, row-block pipeline above 2^18 points) == points handle + bartint_eval_points."""
    rng = np.random.default_rng(77)
    x_train = rng.random((60, 6))
    ff = synth.prior_forest(12, 9, 10, x_train, rng.normal(size=60))
    N = (1 << 18) + 12345
    X = rng.random((N, 6))
    w = rng.random(N)
    f = Forest.from_soa(ff)
    a = f.eval(X, weights=w, want=("draw_mean", "point_mean", "point_var"))            # one-shot, pipelined
    pts = f.points(X)
    b = f.eval(pts, weights=w, want=("draw_mean", "point_mean", "point_var"))          # resident codes
    span = ff.y_max - ff.y_min
    _assert_close(a["draw_mean"], b["draw_mean"], 1e-13, scale=span)
    assert np.array_equal(a["point_mean"], b["point_mean"]) and np.array_equal(a["point_var"], b["point_var"])
    sub = slice(0, N, 997)
    P = onp.predict(ff, X[sub])
    _assert_close(a["point_mean"][sub], P.mean(axis=0), 1e-12, scale=span)
    _assert_close(a["point_var"][sub], onp.col_vars(P) * w[sub], 1e-9, scale=np.maximum(onp.col_vars(P) * w[sub], 1e-12))
    pts.close(); f.close()


def test_staged_matrix_path(small_case):
    from bartint_b200 import StagedMatrix
    ff, _, _, X = small_case
    st = StagedMatrix(X)                    # copy starts before the forest exists
    f = Forest.from_soa(ff)
    pts = st.points(f)
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(X, ff.cut_offsets, ff.cut_values))
    out = f.eval(pts, want=("draw_mean",))
    _assert_close(out["draw_mean"], onp.row_means(onp.predict(ff, X)), 1e-13, scale=ff.y_max - ff.y_min)
    pts.close(); st.close(); f.close()


@pytest.mark.parametrize("S,draws", [(12, None), (12, (4, 12)), (13, None), (12, (3, 11)), (40, (1, 38)), (64, None)])
def test_full_matrix_writer_alignment_cases(S, draws):
    """predict()'s S x N matrix (BARTInt.R:312,380; bartMean.R:47,74): the gather kernel stores four draws of a point as one
    32-byte sector when the row count is a multiple of four, scalars otherwise and at ragged chunk ends; every case must
    be the same matrix, equal to the oracle's."""
    rng = np.random.default_rng(S)
    x_train = rng.random((40, 4))
    ff = synth.prior_forest(S, S, 8, x_train, synth.genz_gaussian(x_train))
    X = rng.random((4099, 4))
    P = onp.predict(ff, X)
    f = Forest.from_soa(ff)
    lo, hi = draws if draws else (0, S)
    out = f.eval(X, want=("full_pred", "draw_mean"), draws=draws)
    _assert_close(out["full_pred"], P[lo:hi], 1e-13, scale=ff.y_max - ff.y_min)
    walk = f.eval(X, want=("full_pred",), draws=draws, force_walk=True)["full_pred"]
    _assert_close(out["full_pred"], walk, 1e-13, scale=ff.y_max - ff.y_min)
    f.close()
