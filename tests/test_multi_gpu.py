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
