"""BASELINE configuration 2 at its FULL size (1000 draws x 50 trees, 10^6 uniform points, 100 candidates) on the GPU:
- against the CPU oracle's values for the same seeded forest and points (tests/golden/bench_cfg2_expected.json, made by
  tools/bench_expected_cpu.py with oracle/oracle_c.c: a minute of host time, so the values are committed), through the
  separate calls (rowMeans(predict) of bartMean.R:74-82, sampleIntegrals + BARTInt.R:260-262, colVars of
  BARTInt.R:312-314) and through the fused host-buffer step bartint_design_step_dbarts098;
- through properties that need no oracle at this size: additivity of the per-draw sums over point shards, the bit-sliced
  kernel against the per-point kernel, draw sub-ranges against the whole, the S x N matrix against the reductions."""
import json
import os

import numpy as np
import pytest

from bartint_b200 import Forest, StagedMatrix, WireStrings, design_step_dbarts, synth
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cfg2_full():
    c = synth.CONFIGS[2]
    x_train, y_train = synth.make_design(2)
    ff = synth.prior_forest(2000, c["S"], c["m"], x_train, y_train)
    X = synth.make_points(2, c["N"], seed=12345)
    Xc = synth.make_points(2, c["C"], seed=777)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_cfg2_expected.json")))["forests"]["2000"]
    return ff, x_train, X, Xc, gold


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_separate_calls_equal_the_oracle_at_full_size(cfg2_full):
    ff, _, X, Xc, gold = cfg2_full
    f = Forest.from_soa(ff)
    assert f.bitslice_ok
    dm = f.eval(X, want=("draw_mean",))["draw_mean"]
    assert f.last_eval_profile(False)["kernel"] == "bitslice"
    assert _rel(float(np.mean(dm)), gold["mc_mean"]) <= 1e-12
    _, imean, isd = onp.finalize_integrals(f.exact_integrals("uniform"), ff.y_min, ff.y_max)      # BARTInt.R:260-262
    assert _rel(float(imean), gold["integral_mean"]) <= 1e-12
    assert _rel(float(isd), gold["integral_sd"]) <= 1e-10
    var = f.eval(Xc, want=("point_var",))["point_var"]
    assert int(np.argmax(var)) == gold["argmax_candidate"]
    f.close()


def test_fused_host_buffer_step_equals_the_oracle_at_full_size(cfg2_full):
    ff, x_train, X, Xc, gold = cfg2_full
    strings, fits = ff.to_dbarts_state(x_train)
    sx, sxc = StagedMatrix(np.asfortranarray(X)), StagedMatrix(np.asfortranarray(Xc))
    o = design_step_dbarts(WireStrings(strings), fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx,
                           candidates=sxc, weights=1.0, measure="uniform")
    sx.close(); sxc.close()
    assert _rel(o["integral_mean_sd"][0], gold["integral_mean"]) <= 1e-12
    assert _rel(o["integral_mean_sd"][1], gold["integral_sd"]) <= 1e-10
    assert _rel(o["draw_mean_var"][0], gold["mc_mean"]) <= 1e-12
    assert int(np.argmax(o["cand_var"])) == gold["argmax_candidate"]


def test_full_size_properties(cfg2_full):
    ff, _, X, _, _ = cfg2_full
    N = len(X)
    f = Forest.from_soa(ff)
    scale = ff.y_max - ff.y_min
    whole = f.eval(X, want=("draw_mean",), scaled=True)["draw_mean"]
    # additivity over ragged point shards (what the multi-GPU layer relies on)
    cuts = [0, 1, 333_333, 700_001, N]
    parts = [f.eval(X[a:b], want=("draw_mean",), scaled=True)["draw_mean"] * (b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.max(np.abs(sum(parts) / N - whole)) <= 1e-13
    # the per-point kernel on the same points
    pp = f.eval(X, want=("draw_mean",), scaled=True, no_bitslice=True)["draw_mean"]
    assert f.last_eval_profile(False)["kernel"] != "bitslice"
    assert np.max(np.abs(pp - whole)) <= 1e-12
    # draw sub-ranges that start and end inside a slice of 32 draws and inside a permutation block
    # (the tile sums are exact integers and the tiles are added in a fixed order: bit-identical to the whole range
    #  evaluated on the same point handle; the one-shot host call above adds row blocks instead, 1 ulp away)
    pts = f.points(X)
    ref = f.eval(pts, want=("draw_mean",), scaled=True, draws=(0, ff.S))["draw_mean"]
    assert np.max(np.abs(ref - whole)) <= 1e-14
    for a, b in [(0, 1), (31, 65), (500, 1000), (999, 1000)]:
        sub = f.eval(pts, want=("draw_mean",), scaled=True, draws=(a, b))["draw_mean"]
        assert np.array_equal(sub, ref[a:b])
    pts.close()
    # the S x N matrix of predict() on a slice of the points against the fused reductions
    sl = X[123_456:123_456 + 4096]
    full = f.eval(sl, want=("full_pred",))["full_pred"]
    red = f.eval(sl, want=("draw_mean", "point_mean", "point_var"))
    assert np.max(np.abs(full.mean(axis=1) - red["draw_mean"])) <= 1e-12 * max(scale, 1.0)
    assert np.max(np.abs(full.mean(axis=0) - red["point_mean"])) <= 1e-12 * max(scale, 1.0)
    assert np.max(np.abs(full.var(axis=0, ddof=1) - red["point_var"])) <= 1e-10 * max(scale * scale, 1.0)
    f.close()


def test_large_host_matrix_arrives_in_eight_row_blocks():
    """A host matrix of 160 MB or more is staged in eight row blocks (four below that): the fused step on 2.2 * 10^6 rows
    equals the separate calls on a point handle."""
    rng = np.random.default_rng(21)
    x_train = rng.random((60, 10))
    ff = synth.prior_forest(77, 48, 12, x_train, synth.genz_gaussian(x_train))
    X = np.asfortranarray(rng.random((2_200_003, 10)))
    Xc = rng.random((50, 10))
    strings, fits = ff.to_dbarts_state(x_train)
    sx, sxc = StagedMatrix(X), StagedMatrix(np.asfortranarray(Xc))
    o = design_step_dbarts(strings, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc, weights=1.0)
    sx.close(); sxc.close()
    f = Forest.from_soa(ff)
    pts = f.points(X)
    want = f.eval(pts, want=("draw_mean",))["draw_mean"]
    pts.close()
    scale = ff.y_max - ff.y_min
    assert np.max(np.abs(o["draw_mean"] - want)) <= 1e-13 * max(scale, 1.0)
    assert np.max(np.abs(o["cand_var"] - f.eval(Xc, want=("point_var",))["point_var"])) <= 1e-10 * max(scale * scale, 1.0)
    f.close()
