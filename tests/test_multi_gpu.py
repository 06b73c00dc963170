"""Multi-GPU parity (run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`): the k-GPU
result of the sharded sequential-design step equals the 1-GPU result.  Skips when fewer than 2 GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from bartint_b200 import Forest, synth, dist as bdist
        rng = np.random.default_rng(21)
        x_train = rng.random((60, 5))
        ff = synth.prior_forest(3, 40, 20, x_train, synth.genz_gaussian(x_train))
        X = rng.random((30011, 5))
        Xc = rng.random((100, 5))
        f = Forest.from_soa(ff)
        lo, hi = bdist.shard_range(len(X), world, rank)
        dX = torch.from_numpy(np.asfortranarray(X[lo:hi]).T.copy()).cuda()
        dXc = torch.from_numpy(np.asfortranarray(Xc).T.copy()).cuda()
        pm = f.points_from_device(dX.data_ptr(), hi - lo)
        pc = f.points_from_device(dXc.data_ptr(), len(Xc))
        dm, mv = bdist.point_sharded_draw_means(f, pm)
        var = bdist.draw_sharded_candidate_var(f, pc, None)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"r{rank}.npy"),
                np.concatenate([dm.cpu().numpy(), mv.cpu().numpy(), var.cpu().numpy()]))
        if rank == 0:     # single-GPU answer on the same inputs
            one = f.eval(X, want=("draw_mean",))["draw_mean"]
            one_var = f.eval(Xc, want=("point_var",))["point_var"]
            np.save(os.path.join(out_dir, "single.npy"), np.concatenate([one, [one.mean(), one.var(ddof=1)], one_var]))
        pm.close(); pc.close(); f.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_k_gpu_equals_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    world = min(_n_gpus(), 8)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npy")
    ranks = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    for r in ranks[1:]:
        assert np.array_equal(r, ranks[0])                 # rank-ordered merges: bit-identical on every rank
    np.testing.assert_allclose(ranks[0], single, rtol=1e-12, atol=1e-15)


def _worker_sharded_step(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from bartint_b200 import Forest, StagedMatrix, WireStrings, design_step_dbarts, synth, dist as bdist
        rng = np.random.default_rng(33)
        x_train = rng.random((70, 6))
        ff = synth.prior_forest(8, 41, 13, x_train, synth.genz_gaussian(x_train))       # 41 draws: ragged shards
        strings, fits = ff.to_dbarts_state(x_train)
        wire = WireStrings(strings)
        X = rng.random((20003, 6))
        Xc = rng.random((77, 6))
        w = rng.random(77)
        lo, hi = bdist.shard_range(len(X), world, rank)
        sx, sc = StagedMatrix(X[lo:hi], rank), StagedMatrix(Xc, rank)
        out = bdist.sharded_design_step(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, sx, sc, weights=w)
        sx.close(); sc.close()
        np.save(os.path.join(out_dir, f"s{rank}.npy"),
                np.concatenate([out[k] for k in ("integrals", "integral_mean_sd", "cand_var", "draw_mean", "draw_mean_var")]))
        # image round trip on one device: an unpacked forest evaluates like its source and refuses export_soa
        f = Forest.from_soa(ff)
        hb, db = f.image_size()
        blob = torch.empty(db, dtype=torch.uint8, device="cuda")
        g = Forest.unpack(f.pack(blob.data_ptr(), db), blob.data_ptr(), db)
        assert (g.S, g.m, g.d, g.n_groups, g.fast_kernel) == (f.S, f.m, f.d, f.n_groups, f.fast_kernel)
        a = f.eval(Xc, want=("full_pred",))["full_pred"]
        b = g.eval(Xc, want=("full_pred",))["full_pred"]
        assert np.array_equal(a, b) and np.array_equal(f.leaf_indices(Xc), g.leaf_indices(Xc))
        assert np.array_equal(f.exact_integrals("uniform"), g.exact_integrals("uniform"))
        try:
            g.export_soa()
            raise AssertionError("export_soa of an unpacked forest must fail")
        except Exception as e:
            assert "device image" in str(e)
        f.close(); g.close()
        if rank == 0:     # single-GPU answer: the fused one-call step on all rows
            sx, sc = StagedMatrix(X, 0), StagedMatrix(Xc, 0)
            one = design_step_dbarts(wire, fits, x_train, ff.cut_lists(), ff.y_min, ff.y_max, mc=sx, candidates=sc, weights=w)
            sx.close(); sc.close()
            np.save(os.path.join(out_dir, "single_step.npy"),
                    np.concatenate([one[k] for k in ("integrals", "integral_mean_sd", "cand_var", "draw_mean", "draw_mean_var")]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_sharded_design_step_equals_one_gpu(tmp_path):
    """Draw-sharded export + all-gathered forest images + point-sharded evaluation == the one-GPU fused step."""
    import torch.multiprocessing as mp
    world = min(_n_gpus(), 8)
    mp.spawn(_worker_sharded_step, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single_step.npy")
    ranks = [np.load(tmp_path / f"s{r}.npy") for r in range(world)]
    for r in ranks[1:]:
        assert np.array_equal(r, ranks[0])
    np.testing.assert_allclose(ranks[0], single, rtol=1e-10, atol=1e-13)
