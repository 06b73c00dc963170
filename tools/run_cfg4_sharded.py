#!/usr/bin/env python
"""BASELINE config 4 as it is worded: "mainRareEvents rare-event integrand d=10, 1e8 Gaussian-measure points sharded
across 8 GPUs" -- 2000 draws x 50 trees (first fit of mainRareEvents.R), the 1e8 points drawn ON THE DEVICE from
N(0.5,1) truncated to [0,1]^10 (the reference's "gaussian" sampler, BARTInt.R:304) with the counter-based Philox
generator, point-sharded: every rank generates and evaluates its own slice [first, first + n) of ONE global stream, the
per-draw sums are all-gathered and added in rank order (bart-int_b200/dist.py).  Launch with torchrun; rank 0 prints
one JSON line.  A parity property rides along: the sharded result must equal a second evaluation in which every rank
re-generates its slice in two halves (same stream, different launch shapes).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_cfg4_sharded.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bartint_b200 import Forest, synth  # noqa: E402
from bartint_b200 import dist as bdist  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    c = synth.CONFIGS[4]
    x_train, y_train = synth.make_design(4)
    ff = synth.prior_forest(4000, c["S"], c["m"], x_train, y_train)
    from bartint_b200 import _lib
    _lib.load().bartint_set_host_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))    # torchrun pins OMP to 1
    f = Forest.from_soa(ff)
    S = ff.S
    lo, hi = bdist.shard_range(n_total, world, rank)
    st = torch.cuda.current_stream().cuda_stream
    seed = 20240404

    def run(pieces):
        sums = torch.zeros(S, dtype=torch.float64, device=dev)
        part = torch.empty(S, dtype=torch.float64, device=dev)
        edges = np.linspace(lo, hi, pieces + 1).astype(np.int64)
        for a, b in zip(edges[:-1], edges[1:]):
            pts = f.generate_points("gaussian", int(b - a), seed, first_index=int(a), stream=st)
            f.eval_device(pts, d_draw_sum=part.data_ptr(), stream=st)
            sums += part
            pts.close()
        total, _ = bdist.allreduce_draw_sums(sums, hi - lo, n_total=n_total)
        return total

    run(1)                                  # warm-up (allocator, NCCL)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    total = run(1)
    b.record(); b.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total2 = run(2)
    torch.cuda.synchronize()
    # dbarts' scaled units (the synthetic rare-event response is constant, so the original-unit range is degenerate)
    draw_mean = total / n_total                                        # rowMeans(predict(model, X)) before the y rescale
    draw_mean2 = total2 / n_total
    same = float((draw_mean - draw_mean2).abs().max())
    if rank == 0:
        t = float(ms.item()) * 1e-3
        print(json.dumps({"cfg": 4, "workload": c["name"], "n_gpus": world, "points_total": n_total,
                          "points_per_gpu": hi - lo, "draws": S, "trees": ff.m, "tree_evals": S * ff.m * n_total,
                          "s": t, "evals_per_s": S * ff.m * n_total / t,
                          "includes": "on-device Philox generation + binning of the shard, ensemble evaluation, all-gather of per-draw sums",
                          "integral_mean_scaled_units": float(draw_mean.mean()),
                          "integral_sd_scaled_units": float(draw_mean.std(unbiased=True)),
                          "max_abs_diff_vs_two_piece_generation": same}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
