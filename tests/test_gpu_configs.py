"""GPU parity at the shapes of the BASELINE configs: >= 1e6 (draw, tree, point) leaf triples bit-equal
to the oracle, and small instances of cfg 1 (d=1), cfg 2, cfg 3 (d=20, 200 trees per draw: the kernels for more
than 16 variables), cfg 4 (gaussian-measure points) and cfg 5 (survey-shaped
integer features with per-candidate variance) against the C restatement of the reference arithmetic."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp
from bartint_b200 import Forest, synth
from bartint_testlib import apply_kernel_kind, kernel_kind_params

pytestmark = pytest.mark.gpu


@pytest.fixture(params=kernel_kind_params(), autouse=True)
def kernel_kind(request, monkeypatch):
    return apply_kernel_kind(request.param, monkeypatch)


def _close(got, want, tol, scale=None):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape
    s = float(np.max(np.abs(want))) if scale is None else scale
    assert np.max(np.abs(got - want)) <= tol * max(s, 1e-300), np.max(np.abs(got - want)) / max(s, 1e-300)


def test_leaf_indices_two_million_triples():
    """cfg 2 shape (d=10, 50 trees): 20 draws x 50 trees x 2000 points = 2e6 triples, a fifth of the points sitting
    exactly on cut values, plus NaN / +-Inf / out-of-range rows."""
    x_train, y_train = synth.make_design(2)
    ff = synth.prior_forest(77, 20, 50, x_train, y_train)
    rng = np.random.default_rng(5)
    X = rng.random((2000, 10))
    cl = ff.cut_lists()
    for j in range(10):
        X[j::50, j] = rng.choice(cl[j], len(X[j::50, j]))
        X[j + 10::50, j] = np.nextafter(rng.choice(cl[j], len(X[j + 10::50, j])), np.inf)
    X[3, 2] = np.nan; X[4, :] = np.inf; X[5, :] = -np.inf; X[6, 7] = -3.0; X[7, 9] = 42.0
    want = oc.leaf_ordinals(ff, X)
    assert want.size >= 1_000_000
    f = Forest.from_soa(ff)
    for table in (False, True):
        assert np.array_equal(f.leaf_indices(X, use_table_kernel=table), want), table
    pts = f.points(X)
    assert np.array_equal(pts.codes().astype(np.int32), onp.bin_points(X, ff.cut_offsets, ff.cut_values))
    pts.close(); f.close()


@pytest.mark.parametrize("cfg,S,N", [(1, 200, 100), (2, 60, 20000), (3, 20, 5000), (4, 40, 20000), (5, 50, 6000)])
def test_config_shapes_small(cfg, S, N):
    c = synth.CONFIGS[cfg]
    x_train, y_train = synth.make_design(cfg, scale=0.2 if cfg == 4 else 1.0)
    ff = synth.prior_forest(1000 * cfg + 1, S, c["m"], x_train, y_train)
    X = synth.make_points(cfg, N)
    w = np.random.default_rng(cfg).random(N)
    ref = oc.predict_reduce(ff, X, nthreads=oc.max_threads(), want=("draw_mean", "point_mean", "point_var"))
    f = Forest.from_soa(ff)
    scale = ff.y_max - ff.y_min
    for walk in (False, True):
        out = f.eval(X, weights=w, want=("draw_mean", "point_mean", "point_var"), force_walk=walk)
        _close(out["draw_mean"], ref["draw_mean"], 1e-12, scale)
        _close(out["point_mean"], ref["point_mean"], 1e-12, scale)
        _close(out["point_var"], ref["point_var"] * w, 1e-9)                      # north-star bar: 1e-6
    # the survey estimator of bartMean.R:74-82 (mean of the per-draw means, their variance without sqrt)
    mean, var = float(np.mean(ref["draw_mean"])), float(np.var(ref["draw_mean"], ddof=1))
    got = f.eval(X, want=("draw_mean",))["draw_mean"]
    assert abs(np.mean(got) - mean) <= 1e-12 * scale and abs(np.var(got, ddof=1) - var) <= 1e-9 * max(var, 1e-300)
    for measure in (("uniform", "gaussian", "exponential") if cfg != 5 else ("uniform",)):
        _close(f.exact_integrals(measure), oc.exact_integrals(ff, measure, nthreads=4), 1e-11, 1.0)
    f.close()
