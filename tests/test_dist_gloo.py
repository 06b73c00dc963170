"""World-size-2 tests of the multi-GPU host layer on CPU (gloo): the sharding, gathering and rank-ordered
merges of bart-int_b200/dist.py, with the oracle standing in for the per-rank kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_np as onp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bartint_b200 import synth, dist as bdist
        rng = np.random.default_rng(seed)
        x_train = rng.random((30, 3))
        ff = synth.prior_forest(seed, 11, 6, x_train, synth.genz_gaussian(x_train))     # S = 11: uneven draw shards
        X = rng.random((1001, 3))                                                         # uneven point shards
        Xc = rng.random((37, 3))
        # ---- point-sharded per-draw means ----
        lo, hi = bdist.shard_range(len(X), world, rank)
        local = torch.from_numpy(onp.tree_sums(ff, X[lo:hi]).sum(axis=1))
        total, n_total = bdist.allreduce_draw_sums(local, hi - lo)
        assert n_total == len(X)
        want = onp.tree_sums(ff, X).sum(axis=1)
        np.testing.assert_allclose(total.numpy(), want, rtol=1e-13, atol=1e-13)
        # ---- draw-sharded candidate moments ----
        dlo, dhi = bdist.shard_range(ff.S, world, rank)
        F = onp.tree_sums(ff, Xc)[dlo:dhi]
        mean = torch.from_numpy(F.mean(axis=0))
        m2 = torch.from_numpy(((F - F.mean(axis=0)) ** 2).sum(axis=0))
        counts, means, m2s = bdist.gather_moments(mean, m2, dhi - dlo)
        assert counts.tolist() == [bdist.shard_range(ff.S, world, r)[1] - bdist.shard_range(ff.S, world, r)[0]
                                   for r in range(world)]
        n, mu, M2 = bdist.chan_merge_np(counts, means.numpy(), m2s.numpy())
        assert n == ff.S
        scale = ff.y_max - ff.y_min
        np.testing.assert_allclose(M2 / (n - 1) * scale ** 2, onp.col_vars(onp.predict(ff, Xc)), rtol=1e-11)
        np.testing.assert_allclose(scale * (mu + 0.5) + ff.y_min, onp.predict(ff, Xc).mean(axis=0), rtol=1e-12)
        # every rank must hold bit-identical merged results (rank-ordered, deterministic)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([total.numpy(), M2]))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from bartint_b200.dist import shard_range
    for total in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharding_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 5, str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(tmp_path / f"r{r}.npy") for r in range(world))
    assert np.array_equal(a, b)
