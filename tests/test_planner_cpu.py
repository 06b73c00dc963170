"""CPU tests of the host-side group planner (forest_host.cu: plan_fast_path), through the host-only hook
bartint_debug_plan_soa and the numpy record interpreter in plan_interp.py: for every kernel format, number of
variables and counting mode the records must reproduce the oracle's predictions exactly as the device would form
them (same selectors, biases, weights, tables and flags)."""
import numpy as np
import pytest

import plan_interp as pi
from bartint_b200 import synth
from oracle import oracle_np as onp


def _case(d, seed, S=4, m=24, N=257, numcut=100):
    rng = np.random.default_rng(seed)
    x_train = rng.random((60, d))
    ff = synth.prior_forest(seed, S, m, x_train, numcut=numcut)
    X = rng.random((N, d))
    X[:3] = [0.0, 1.0, 0.5][: 3] if d == 0 else np.array([[0.0] * d, [1.0] * d, [0.5] * d])
    codes = onp.bin_points(X, ff.cut_offsets, ff.cut_values)
    return ff, X, codes


def _check_against_oracle(ff, X, codes, monkeypatch, kernel, count):
    monkeypatch.setenv("BARTINT_KERNEL", kernel)
    monkeypatch.setenv("BARTINT_COUNT_SPLITS", str(count))
    plan = pi.debug_plan(ff)
    assert plan["table_ok"]
    assert plan["gather"] == (kernel == "gather" and ff.d <= 16)
    assert plan["count_level"] == (count if plan["gather"] else 0)
    pi.check_structure(plan)
    want = onp.tree_sums(ff, X)
    # evaluation with per-point results: every record evaluated, draws closed at END_DRAW
    got, skipped = pi.interpret(plan, codes, sums_only=False)
    assert not any(skipped)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
    # sums-only evaluation: COUNTABLE records skipped; their trees are exactly the ones the counting kernel adds
    got, skipped = pi.interpret(plan, codes, sums_only=True)
    csum, counted = pi.count_sums(plan, codes) if plan["count_level"] else (np.zeros(ff.S), [[] for _ in range(ff.S)])
    nodes = np.diff(np.asarray(ff.tree_offset))
    for s in range(ff.S):
        # the trees the planner put behind END_EVAL are exactly the trees the counting kernel adds
        assert sorted(skipped[s]) == sorted(counted[s])
        ks = np.arange(s * ff.m, (s + 1) * ff.m)
        small = {int(k) for k in ks if 3 <= nodes[k] <= 1 + 2 * plan["count_level"]}
        assert set(skipped[s]) <= small
        if ff.d <= 12:
            assert set(skipped[s]) == small               # up to three code words every small tree fits a DUAL table
    # per-draw sums: evaluated records + histogram counts == oracle
    np.testing.assert_allclose(got.sum(axis=1) + csum, want.sum(axis=1), rtol=1e-12, atol=1e-12)
    for s in range(ff.S):
        for k in skipped[s]:
            got[s] += pi.walk_tree(plan, k, codes)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
    return plan


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 8, 9, 10, 12, 13, 14, 16])
@pytest.mark.parametrize("count", [0, 1, 2])
def test_gather_records_reproduce_the_oracle(d, count, monkeypatch):
    ff, X, codes = _case(d, seed=100 + d)
    plan = _check_against_oracle(ff, X, codes, monkeypatch, "gather", count)
    assert plan["regs"] == (2 if d <= 8 else 3 if d <= 12 else 4)


@pytest.mark.parametrize("d", [1, 7, 10, 20])
def test_table_records_reproduce_the_oracle(d, monkeypatch):
    ff, X, codes = _case(d, seed=200 + d, m=16)
    _check_against_oracle(ff, X, codes, monkeypatch, "table", 0)


@pytest.mark.parametrize("bits", [4, 5, 8])
def test_table_bits(bits, monkeypatch):
    monkeypatch.setenv("BARTINT_TABLE_BITS", str(bits))
    ff, X, codes = _case(6, seed=300 + bits, S=2, m=12, N=65)
    plan = _check_against_oracle(ff, X, codes, monkeypatch, "table", 0)
    assert plan["nb"] == bits
    plan = _check_against_oracle(ff, X, codes, monkeypatch, "gather", 0)
    assert plan["nb"] == min(bits, 5)


@pytest.mark.parametrize("kind", ["stumps", "left_chain", "right_chain", "same_var", "balanced", "edge_cuts"])
@pytest.mark.parametrize("d", [1, 3, 11, 16])
@pytest.mark.parametrize("count", [0, 1, 2])
def test_adversarial_shapes(kind, d, count, monkeypatch):
    # deep chains become WALK records, 5-node trees JOINT records, same_var trees put all nodes on one code word
    ff = synth.adversarial_forest(kind, 2, 7, d, numcut=100, seed=d, depth=7)
    rng = np.random.default_rng(d)
    X = rng.random((130, d))
    codes = onp.bin_points(X, ff.cut_offsets, ff.cut_values)
    plan = _check_against_oracle(ff, X, codes, monkeypatch, "gather", count)
    flags = plan["hdr"]
    if kind in ("left_chain", "right_chain"):
        assert (flags & pi.FLAG_WALK).any()
    if kind == "stumps":
        assert plan["n_groups"] == ff.S                      # one constant record per draw


def test_five_node_trees_use_joint_records(monkeypatch):
    cl = [np.linspace(0, 1, 102)[1:-1]] * 10
    five = (0, 10, (1, 20, 0.01, 0.02), (5, 30, (9, 40, 0.03, 0.04), (2, 50, 0.05, 0.06)))     # 5 internal nodes
    ff = synth.forest_from_trees([five, 0.5, (3, 3, -0.1, 0.1)] * 2, 2, 3, cl)
    rng = np.random.default_rng(0)
    X = rng.random((99, 10))
    codes = onp.bin_points(X, ff.cut_offsets, ff.cut_values)
    plan = _check_against_oracle(ff, X, codes, monkeypatch, "gather", 0)
    assert (plan["hdr"] & pi.FLAG_JOINT).any()


def test_planner_is_deterministic_and_thread_count_independent(monkeypatch):
    import ctypes as C
    from bartint_b200 import _lib
    monkeypatch.setenv("BARTINT_KERNEL", "gather")
    monkeypatch.setenv("BARTINT_COUNT_SPLITS", "0")
    ff, _, _ = _case(10, seed=7, S=16, m=50)
    lib = _lib.load()
    before = lib.bartint_set_host_threads(1)
    a = pi.debug_plan(ff)
    lib.bartint_set_host_threads(4)
    b = pi.debug_plan(ff)
    lib.bartint_set_host_threads(0)
    for k in ("hdr", "split", "desc", "tree_perm", "group_tree_off", "draw_group_off", "node_bit"):
        assert np.array_equal(a[k], b[k]), k
    # bin packing reaches the lower bound of ceil(internal nodes / 8) DUAL records for almost every draw
    nodes = np.diff(np.asarray(ff.tree_offset))
    internal = ((nodes - 1) // 2).reshape(ff.S, ff.m).sum(axis=1)
    recs = np.diff(a["draw_group_off"])
    assert (recs >= np.ceil(internal / 8)).all()
    assert recs.sum() <= np.ceil(internal / 8).sum() + ff.S


def test_capacity_too_small_is_an_error_not_an_overrun():
    import ctypes as C
    from bartint_b200 import _lib
    ff, _, _ = _case(4, seed=1, S=3, m=20)
    lib = _lib.load()
    info = np.zeros(8, np.int64)
    toff = np.ascontiguousarray(ff.tree_offset, np.int64)
    a32 = lambda a: np.ascontiguousarray(a, np.int32)
    var, cut, left, right, coff = a32(ff.var), a32(ff.cut), a32(ff.left), a32(ff.right), a32(ff.cut_offsets)
    mu, cval = np.ascontiguousarray(ff.mu, np.float64), np.ascontiguousarray(ff.cut_values, np.float64)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.bartint_debug_plan_soa(ff.S, ff.m, ff.d, p(toff, C.c_int64), p(var, C.c_int32), p(cut, C.c_int32),
                                    p(left, C.c_int32), p(right, C.c_int32), p(mu, C.c_double), p(coff, C.c_int32),
                                    p(cval, C.c_double), 1, p(info, C.c_int64), None, None, None, None, None, None,
                                    None, None, None, None)
    assert rc == 1 and b"capacity" in lib.bartint_last_error()
    assert info[5] > 1


def test_two_split_trees_on_unpackable_words_stay_evaluated(monkeypatch):
    """Four code words (13 <= d <= 16): a two-split tree on words (0,2) or (1,2) fits no DUAL table.  It must stay in
    the evaluated part (the unmerged level-2 patch looped forever here) and must not be counted."""
    cl = [np.linspace(0, 1, 102)[1:-1]] * 16
    trees = [(0, 10, 0.1, (9, 20, 0.2, 0.3)),      # words (0,2): not DUAL-able
             (5, 30, (10, 40, -0.1, -0.2), 0.05),  # words (1,2): not DUAL-able
             (1, 50, 0.3, (14, 60, 0.1, 0.0)),     # words (0,3): table 2, pair (0,R-1)
             (8, 70, (13, 5, 0.2, 0.1), -0.3),     # words (2,3): table 2, pair (R-2,R-1)
             (2, 15, 0.0, (6, 25, 0.4, 0.5)),      # words (0,1): table 1
             (11, 99, 0.7, 0.8)]                   # single split on word 2
    ff = synth.forest_from_trees(trees * 3, 3, 6, cl)
    rng = np.random.default_rng(5)
    X = rng.random((333, 16))
    codes = onp.bin_points(X, ff.cut_offsets, ff.cut_values)
    plan = _check_against_oracle(ff, X, codes, monkeypatch, "gather", 2)
    _, skipped = pi.interpret(plan, codes, sums_only=True)
    assert [sorted(k % 6 for k in sk) for sk in skipped] == [[2, 3, 4, 5]] * 3


def test_plan_digests_match_the_gpu_verified_build():
    """tests/golden/plan_digests.json pins what the planner uploads (tools/plan_digest.py: 90 forests x formats x
    counting levels).  The level-0 digests have not changed since the host-only hook exists (the planner body was moved
    out of forest_upload verbatim from the build verified on the GPU), and every one of these plans reproduces the
    oracle through the interpreter above.  A change that alters the digests must be deliberate: re-verify on a GPU,
    then regenerate the file with `python tools/plan_digest.py > tests/golden/plan_digests.json`."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if not k.startswith("BARTINT_")}
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "plan_digest.py")], check=True,
                         capture_output=True, text=True, env=env).stdout
    with open(os.path.join(root, "tests", "golden", "plan_digests.json")) as fh:
        want = json.load(fh)
    got = json.loads(out)
    assert sorted(got) == sorted(want)
    changed = [k for k in want if got[k] != want[k]]
    assert not changed, f"planner output changed for {changed[:8]} ..."


def test_randomized_sweep(monkeypatch):
    """Seeded sweep over d, trees per draw, cut grid sizes, table bits, formats, counting levels and tree shapes."""
    rng = np.random.default_rng(12345)
    for it in range(90):
        d = int(rng.integers(1, 21))
        m = int(rng.integers(1, 60))
        S = int(rng.integers(1, 4))
        numcut = int(rng.choice([1, 2, 5, 100, 127]))
        kernel = "table" if d > 16 or rng.random() < 0.3 else "gather"
        count = int(rng.integers(0, 3)) if kernel == "gather" else 0
        monkeypatch.setenv("BARTINT_TABLE_BITS", str(int(rng.integers(4, 9))))
        x = rng.random((int(rng.integers(5, 80)), d))
        kind = rng.choice(["prior", "prior", "prior", "balanced", "same_var", "left_chain", "edge_cuts"])
        if kind == "prior":
            ff = synth.prior_forest(int(rng.integers(1 << 30)), S, m, x, numcut=numcut, k=float(rng.choice([2.0, 0.5])))
        else:
            ff = synth.adversarial_forest(kind, S, min(m, 12), d, numcut=max(numcut, 2), seed=it,
                                          depth=int(rng.integers(1, 8)))
        X = rng.random((97, d))
        codes = onp.bin_points(X, ff.cut_offsets, ff.cut_values)
        _check_against_oracle(ff, X, codes, monkeypatch, kernel, count)


def test_non_preorder_input_is_canonicalised(monkeypatch):
    """bartint_forest_from_soa accepts any node order (children by index, root first); the library re-lays every tree
    out in pre-order before planning.  Shuffle the nodes of every tree and compare with the oracle on the shuffled input."""
    from bartint_b200 import FlatForest
    rng = np.random.default_rng(21)
    x = rng.random((50, 9))
    ff = synth.prior_forest(21, 3, 30, x)
    off = np.asarray(ff.tree_offset)
    var, cut, left, right, mu = (np.array(getattr(ff, k)) for k in ("var", "cut", "left", "right", "mu"))
    for k in range(ff.S * ff.m):
        a, b = off[k], off[k + 1]
        cnt = b - a
        if cnt < 4:
            continue
        new_of_old = np.concatenate([[0], 1 + rng.permutation(cnt - 1)])         # root stays first
        old_of_new = np.argsort(new_of_old)
        remap = lambda child: np.where(child >= 0, new_of_old[np.maximum(child, 0)], -1)
        v, c, l, r, u = var[a:b][old_of_new], cut[a:b][old_of_new], left[a:b][old_of_new], right[a:b][old_of_new], mu[a:b][old_of_new]
        var[a:b], cut[a:b], left[a:b], right[a:b], mu[a:b] = v, c, remap(l), remap(r), u
    shuffled = FlatForest(ff.S, ff.m, ff.d, off, var, cut, left, right, mu, ff.cut_offsets, ff.cut_values, ff.y_min, ff.y_max)
    X = rng.random((150, 9))
    np.testing.assert_allclose(onp.tree_sums(shuffled, X), onp.tree_sums(ff, X), rtol=0, atol=1e-15)   # same forest
    codes = onp.bin_points(X, ff.cut_offsets, ff.cut_values)
    monkeypatch.setenv("BARTINT_KERNEL", "gather")
    monkeypatch.setenv("BARTINT_COUNT_SPLITS", "0")
    a_plan, b_plan = pi.debug_plan(shuffled), pi.debug_plan(ff)
    for key in ("hdr", "split", "desc", "tree_perm", "node_bit", "walk_nodes", "leaf_mu"):
        assert np.array_equal(a_plan[key], b_plan[key]), key                      # identical after canonicalisation
    got, _ = pi.interpret(a_plan, codes, sums_only=False)
    np.testing.assert_allclose(got, onp.tree_sums(ff, X), rtol=0, atol=1e-13)
