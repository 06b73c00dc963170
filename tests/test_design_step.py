"""bartint_design_step_dbarts098: the fused, draw-chunk-pipelined sequential-design step must give what the separate
ABI calls give (and therefore what the oracle gives), for every chunk count."""
import numpy as np
import pytest

from oracle import oracle_np as onp
from oracle import oracle_c as oc
from bartint_b200 import Forest, StagedMatrix, WireStrings, design_step_dbarts, synth, BartIntError
from bartint_testlib import apply_kernel_kind, kernel_kind_params

pytestmark = pytest.mark.gpu


@pytest.fixture(params=kernel_kind_params(), autouse=True)
def kernel_kind(request, monkeypatch):
    return apply_kernel_kind(request.param, monkeypatch)


def _case(S=37, m=11, d=6, n=80, N=5000, C=300, seed=3):
    rng = np.random.default_rng(seed)
    x_train = rng.random((n, d))
    y_train = synth.genz_gaussian(x_train)
    ff = synth.prior_forest(seed, S, m, x_train, y_train)
    strings, fits = ff.to_dbarts_state(x_train)
    return ff, x_train, strings, fits, rng.random((N, d)), rng.random((C, d)), rng.random(C)


@pytest.mark.parametrize("chunks", [0, 1, 2, 5, 37])
def test_design_step_matches_oracle(chunks):
    ff, x_train, strings, fits, X, Xc, w = _case()
    mc, cand = StagedMatrix(X), StagedMatrix(Xc)
    out = design_step_dbarts(WireStrings(strings), fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=mc, candidates=cand,
                             weights=w, measure="uniform", n_chunks=chunks)
    mc.close(); cand.close()
    scale = ff.y_max - ff.y_min
    integ = oc.exact_integrals(ff, "uniform", nthreads=2)
    assert np.max(np.abs(out["integrals"] - integ)) <= 1e-13
    _, imean, isd = onp.finalize_integrals(integ, ff.y_min, ff.y_max)
    assert abs(out["integral_mean_sd"][0] - imean) <= 1e-12 * scale and abs(out["integral_mean_sd"][1] - isd) <= 1e-10 * isd
    ref = oc.predict_reduce(ff, X, nthreads=4, want=("draw_mean",))
    assert np.max(np.abs(out["draw_mean"] - ref["draw_mean"])) <= 1e-12 * scale
    assert abs(out["draw_mean_var"][0] - np.mean(ref["draw_mean"])) <= 1e-12 * scale
    v = np.var(ref["draw_mean"], ddof=1)
    assert abs(out["draw_mean_var"][1] - v) <= 1e-9 * v
    cref = oc.predict_reduce(ff, Xc, nthreads=4, want=("point_var",))["point_var"] * w
    assert np.max(np.abs(out["cand_var"] - cref)) <= 1e-9 * np.max(cref)      # north-star bar: 1e-6


def test_design_step_equals_separate_calls_and_options():
    ff, x_train, strings, fits, X, Xc, w = _case(S=300, m=7, N=3000, C=64, seed=9)
    wire = WireStrings(strings)
    f = Forest.from_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max)
    sep_var = f.eval(Xc, weights=np.array([2.5]), want=("point_var",))["point_var"]
    sep_dm = f.eval(X, want=("draw_mean",))["draw_mean"]
    sep_int = f.exact_integrals("exponential", n_draws_used=7)
    f.close()
    mc, cand = StagedMatrix(X), StagedMatrix(Xc)
    out = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=mc, candidates=cand,
                             weights=2.5, measure="exponential", n_draws_integral=7)
    assert out["integrals"].shape == (7,) and np.array_equal(out["integrals"], sep_int)
    assert np.max(np.abs(out["draw_mean"] - sep_dm)) <= 1e-13 * (ff.y_max - ff.y_min)
    assert np.max(np.abs(out["cand_var"] - sep_var)) <= 1e-10 * np.max(sep_var)
    # only candidates / only MC points
    o2 = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, candidates=cand, weights=2.5)
    assert set(o2) == {"integrals", "integral_mean_sd", "cand_var"} and np.allclose(o2["cand_var"], sep_var, rtol=1e-10)
    o3 = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=mc, want=("draw_mean",))
    assert set(o3) == {"draw_mean"} and np.allclose(o3["draw_mean"], sep_dm, rtol=0, atol=1e-13)
    mc.close(); cand.close()


def test_design_step_errors():
    ff, x_train, strings, fits, X, Xc, w = _case(S=4, m=3, N=50, C=10)
    cand = StagedMatrix(Xc)
    with pytest.raises(BartIntError):          # weights of the wrong length
        design_step_dbarts(WireStrings(strings), fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, candidates=cand,
                           weights=np.ones(3))
    bad = np.array(strings, dtype=object).copy()
    bad[1, 2] = "0 1 . "                        # truncated tree in the third draw
    with pytest.raises(BartIntError):
        design_step_dbarts(WireStrings(bad), fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, candidates=cand, n_chunks=4)
    wrong_d = StagedMatrix(np.zeros((5, ff.d + 1)))
    with pytest.raises(BartIntError):
        design_step_dbarts(WireStrings(strings), fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=wrong_d)
    wrong_d.close(); cand.close()


def test_api_design_step_matches_drop_in_functions():
    from bartint_b200 import api
    rng = np.random.default_rng(21)
    x = rng.random((60, 3))
    y = synth.genz_gaussian(x)
    model = api.synthetic_bart(x, y, ndpost=400, keepevery=1, ntree=20, seed=4)
    cand = rng.random((100, 3))
    w = rng.random(100)
    step = api.bartint_design_step(model, cand, w, "uniform")
    integ = api.sampleIntegrals(model, 3, "uniform")                      # first ntree draws (quirk B1)
    assert step["integrals"].shape == integ.shape and np.array_equal(step["integrals"], integ)
    resc = (integ + 0.5) * (model.y_max - model.y_min) + model.y_min       # BARTInt.R:260-262
    assert abs(step["integral_mean_sd"][0] - resc.mean()) <= 1e-12 * (model.y_max - model.y_min)
    assert abs(step["integral_mean_sd"][1] - resc.std(ddof=1)) <= 1e-10 * resc.std(ddof=1)
    var = api.bartint_candidate_var(model, cand, w)
    assert np.max(np.abs(step["cand_var"] - var)) <= 1e-10 * np.max(var)
    model.close()


@pytest.mark.parametrize("N", [262144, 300001])
def test_design_step_row_block_path(N):
    """>= 2^18 rows: the matrix is staged in four row blocks and the first chunk of draws is evaluated block by block on
    views of the point set (ragged last block included)."""
    ff, x_train, strings, fits, _, Xc, w = _case(S=24, m=6, d=5, N=10, C=50, seed=17)
    rng = np.random.default_rng(N)
    X = rng.random((N, 5))
    X[::1000, 2] = np.nan
    mc, cand = StagedMatrix(X), StagedMatrix(Xc)
    out = design_step_dbarts(WireStrings(strings), fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=mc, candidates=cand,
                             weights=w, n_chunks=3)
    # the staged handle stays usable for an ordinary consumer afterwards
    f = Forest.from_soa(ff)
    pts = mc.points(f)
    sep = f.eval(pts, want=("draw_mean",))["draw_mean"]
    pts.close(); f.close(); mc.close(); cand.close()
    ref = oc.predict_reduce(ff, X, nthreads=oc.max_threads(), want=("draw_mean",))["draw_mean"]
    scale = ff.y_max - ff.y_min
    assert np.max(np.abs(out["draw_mean"] - ref)) <= 1e-12 * scale
    assert np.max(np.abs(out["draw_mean"] - sep)) <= 1e-13 * scale


@pytest.mark.gpu
def test_topk_batch_selection():
    """order(var, decreasing = TRUE)[1:k] on the device: ties in index order, NaN never selected, k beyond the data -> -1."""
    from bartint_b200 import topk
    rng = np.random.default_rng(0)
    for n, k in [(100, 1), (100, 7), (5000, 64), (200_000, 20), (3, 5)]:
        v = np.round(rng.random(n), 2)                 # many ties
        if n > 10:
            v[rng.integers(0, n, 5)] = np.nan
        idx, val = topk(v, k)
        order = sorted((i for i in range(n) if v[i] == v[i]), key=lambda i: (-v[i], i))
        want = (order + [-1] * k)[:k]
        assert list(idx) == want, (n, k)
        assert all((val[r] == v[want[r]]) if want[r] >= 0 else np.isnan(val[r]) for r in range(k))
    assert list(topk(np.array([np.nan, np.nan]), 2)[0]) == [-1, -1]
    assert len(topk(np.zeros(4), 0)[0]) == 0
