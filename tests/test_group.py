"""The single-process device group behind the C ABI (bartint_init, include/bartint.h): one call drives every GPU of
the box, which is what the reference's one R process can reach (src/BARTInt.R:259-262, 312-314; bartMean.R:74-82).
The k-member result of bartint_design_step_dbarts098 must equal the one-GPU result.  On a box with one GPU the
group is opened over the same device several times (BARTINT_GROUP_DEVICES, peer-copy exchange); with >= 2 GPUs the
members are distinct devices and the exchange is ncclAllGather."""
import os

import numpy as np
import pytest

import bartint_b200 as bi
from bartint_b200 import StagedMatrix, design_step_dbarts, synth

pytestmark = pytest.mark.gpu


def _n_gpus():
    import ctypes as C
    from bartint_b200 import _lib
    n = C.c_int32(0)
    _lib.load().bartint_device_count(C.byref(n))
    return n.value


def _case(S=96, m=12, d=4, N=70_001, C=64, seed=3):
    rng = np.random.default_rng(seed)
    x_train = rng.random((50, d))
    y = synth.genz_gaussian(x_train)
    ff = synth.prior_forest(seed, S, m, x_train, y)
    strings, fits = ff.to_dbarts_state(x_train)
    return ff, x_train, strings, fits, rng.random((N, d)), rng.random((C, d))


def _step(case, device, n_chunks=0):
    ff, x_train, strings, fits, X, Xc = case
    sx, sxc = StagedMatrix(X, device), StagedMatrix(Xc, 0)
    try:
        return design_step_dbarts(strings, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sxc,
                                  weights=1.0, measure="uniform", n_chunks=n_chunks)
    finally:
        sx.close(); sxc.close()


def _same(a, b, tol=1e-12):
    for k in ("integrals", "integral_mean_sd", "cand_var", "draw_mean", "draw_mean_var"):
        s = max(float(np.max(np.abs(a[k]))), 1e-300)
        assert np.max(np.abs(np.asarray(a[k]) - np.asarray(b[k]))) <= tol * s, k


@pytest.mark.parametrize("members", [1, 2, 3])
def test_group_on_one_device_equals_single(members, monkeypatch):
    """Members on the SAME device (test mode): worker threads, image replication, sharded staging and the
    member-ordered exchange are exercised on a one-GPU box."""
    case = _case()
    want = _step(case, 0)
    monkeypatch.setenv("BARTINT_GROUP_DEVICES", ",".join(["0"] * members))
    try:
        assert bi.init_group(members) == members
        for n_chunks in (0, 1, 3):
            got = _step(case, "group", n_chunks)
            _same(got, want)
        # a matrix with fewer rows than members x 1024: trailing members hold empty shards
        small = _case(N=1500)
        _same(_step(small, "group"), _step(small, 0))
    finally:
        bi.shutdown_group()
    assert bi.group_size() == 1


def test_group_shards_in_row_blocks(monkeypatch):
    """Shards of >= 2^18 rows arrive in four row blocks; every member bins and evaluates its shard block by block (the
    last member's shard is shorter and ragged), for one and for several chunks of draws."""
    case = _case(S=40, m=8, d=3, N=2 * 300_032 + 270_001, C=16, seed=11)
    want = _step(case, 0)
    monkeypatch.setenv("BARTINT_GROUP_DEVICES", "0,0,0")
    try:
        assert bi.init_group(3) == 3
        for n_chunks in (0, 1, 3):
            _same(_step(case, "group", n_chunks), want)
    finally:
        bi.shutdown_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_group_over_all_gpus_equals_single(monkeypatch):
    case = _case(S=200, m=20, d=6, N=300_007, C=100, seed=8)
    big = _case(S=64, m=10, d=4, N=_n_gpus() * 270_336 - 5, C=32, seed=9)      # shards that arrive in row blocks
    want, want_big = _step(case, 0), _step(big, 0)
    try:
        g = bi.init_group(0)
        assert g == _n_gpus()
        # the gather-to-root of the step: peer stores into member 0's buffer (default) and the NCCL exchange
        assert bi.group_exchange().startswith("peer writes")
        for _ in range(2):
            _same(_step(case, "group"), want)
        _same(_step(big, "group"), want_big)
        monkeypatch.setenv("BARTINT_GROUP_EXCHANGE", "nccl")
        assert bi.group_exchange().startswith("ncclAllGather")
        _same(_step(case, "group"), want)
        _same(_step(big, "group", 2), want_big)
        monkeypatch.delenv("BARTINT_GROUP_EXCHANGE")
        assert bi.init_group(2) == 2          # re-opening with another size replaces the group
        _same(_step(case, "group"), want)
    finally:
        bi.shutdown_group()


def _eval_case(seed=5, S=70, m=9, d=4, N=9000):
    rng = np.random.default_rng(seed)
    x_train = rng.random((40, d))
    ff = synth.prior_forest(seed, S, m, x_train, synth.genz_gaussian(x_train))
    return ff, rng.random((N, d)), rng.random(N)


def _eval_all(ff, X, w):
    """predict / colVars * weights / rowMeans through the one-shot host entry point (what the R shim calls)"""
    from bartint_b200 import Forest
    f = Forest.from_soa(ff)
    try:
        a = f.eval(X, weights=w, want=("draw_mean", "point_mean", "point_var"))
        b = f.eval(X, want=("draw_mean",))
        c = f.eval(X, want=("full_pred", "draw_mean"))
        return a, b, c
    finally:
        f.close()


@pytest.mark.parametrize("members", [2, 3])
def test_group_eval_equals_single(members, monkeypatch):
    """bartint_eval on the device group: rows sharded over the members (forest replicas pulled from member 0), per-point
    outputs written in place, per-draw sums exchanged -- equal to the one-GPU call."""
    ff, X, w = _eval_case()
    want = _eval_all(ff, X, w)
    monkeypatch.setenv("BARTINT_GROUP_DEVICES", ",".join(["0"] * members) if _n_gpus() < members else ",".join(map(str, range(members))))
    monkeypatch.setenv("BARTINT_GROUP_MIN_ROWS", "1024")
    try:
        assert bi.init_group(members) == members
        for rep in range(2):                      # the second pass reuses the cached replicas
            got = _eval_all(ff, X, w)
            for g, wnt in zip(got, want):
                for k in wnt:
                    scale = max(float(np.max(np.abs(wnt[k]))), 1e-300)
                    assert np.max(np.abs(g[k] - wnt[k])) <= 1e-12 * scale, k
    finally:
        bi.shutdown_group()
