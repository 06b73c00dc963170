"""Multi-GPU host layer: one process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in
the CPU tests).  The path shards two ways:

  point-sharded  every rank holds the whole forest and N/G points; the only exchange is the per-draw
                 partial sums (S doubles): all-gather + fixed-order sum (a deterministic all-reduce).
  draw-sharded   every rank holds S/G draws and all candidate points; per-draw results are complete
                 locally (all-gather), per-candidate moments (count, mean, M2) are all-gathered and
                 Chan-merged in rank order.

Payloads are KB-sized, so the exchange is latency-bound; determinism (rank-ordered merges, no
floating-point atomics) is what matters: k-GPU results equal 1-GPU results to ~1e-15.
The local compute and the final merge are injected so the same functions run on CPU tensors in the
gloo tests (with the oracle as the local evaluator) and on CUDA tensors in production.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced split of range(total): the first total % world ranks get one extra item."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def world_info(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """[world, *t.shape] with row r = rank r's tensor (all ranks must pass the same shape)."""
    rank, world = world_info(group)
    if world == 1:
        return t.unsqueeze(0)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group) if t.is_cuda else \
        dist.all_gather(list(out.unbind(0)), t.contiguous(), group=group)
    return out


def ordered_sum(rows: torch.Tensor) -> torch.Tensor:
    """Sum over dim 0 in index order (deterministic, identical on every rank)."""
    acc = rows[0].clone()
    for r in range(1, rows.shape[0]):
        acc += rows[r]
    return acc


def allreduce_draw_sums(local_sums: torch.Tensor, n_local: int, group=None, n_total: int | None = None):
    """Point-sharded exchange: per-draw sums of every rank's points -> (total sums, total point count).

    Pass n_total when the caller knows the global point count: gathering it costs a device->host sync, which would
    stall the launch queue behind the evaluation kernel."""
    rows = all_gather_rows(local_sums, group)
    if n_total is None:
        rank, world = world_info(group)
        if world == 1:
            n_total = n_local
        else:
            counts = all_gather_rows(torch.tensor([n_local], dtype=torch.int64, device=local_sums.device), group)
            n_total = int(counts.sum().item())
    return (ordered_sum(rows) if rows.shape[0] > 1 else rows[0]), n_total


def gather_moments(local_mean: torch.Tensor, local_m2: torch.Tensor, n_local_draws: int, group=None, counts=None):
    """Draw-sharded exchange: (draw counts per rank, means [world, C], M2s [world, C])."""
    means = all_gather_rows(local_mean, group)
    m2s = all_gather_rows(local_m2, group)
    if counts is None:
        c = all_gather_rows(torch.tensor([n_local_draws], dtype=torch.int32, device=local_mean.device), group)
        counts = c.flatten().cpu().numpy().astype(np.int32)          # device->host sync; pass counts to avoid it
    return np.asarray(counts, np.int32), means, m2s


def chan_merge_np(counts, means, m2s):
    """Reference (numpy) rank-ordered Chan merge of (n, mean, M2); mirrors merge_shards_kernel."""
    n = 0.0
    mean = np.zeros(means.shape[1])
    m2 = np.zeros(means.shape[1])
    for g in range(len(counts)):
        nb = float(counts[g])
        if nb == 0:
            continue
        if n == 0:
            n, mean, m2 = nb, means[g].copy(), m2s[g].copy()
            continue
        tot = n + nb
        delta = means[g] - mean
        mean = mean + delta * (nb / tot)
        m2 = m2 + m2s[g] + delta * delta * (n * nb / tot)
        n = tot
    return n, mean, m2


# ---------------------------------------------------------------------------------------------
# production wrappers (CUDA tensors + libbartint_cuda)
# ---------------------------------------------------------------------------------------------
def point_sharded_draw_means(forest, pts_local, group=None, scaled=False, stream=None, n_total=None):
    """rowMeans(predict(model, X)) with X's rows sharded across ranks.  Returns (draw_mean[S], mean_var[2]) on device."""
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    sums = torch.empty(forest.S, dtype=torch.float64, device=dev)
    forest.eval_device(pts_local, d_draw_sum=sums.data_ptr(), stream=st)
    total, n_total = allreduce_draw_sums(sums, pts_local.N, group, n_total)
    out = torch.empty(forest.S, dtype=torch.float64, device=dev)
    mv = torch.empty(2, dtype=torch.float64, device=dev)
    forest.finalize_draws_device(total.data_ptr(), forest.S, n_total, out.data_ptr(), mv.data_ptr(), scaled=scaled, stream=st)
    return out, mv


def draw_sharded_candidate_var(forest, pts, weights=None, group=None, scaled=False, stream=None):
    """colVars(predict(model, X)) * weights with the DRAWS sharded across ranks (each rank evaluates its own
    contiguous range of the forest's draws on all candidates).  Returns var[C] on device."""
    rank, world = world_info(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    lo, hi = shard_range(forest.S, world, rank)
    mean = torch.zeros(pts.N, dtype=torch.float64, device=dev)
    m2 = torch.zeros(pts.N, dtype=torch.float64, device=dev)
    if hi > lo:
        forest.eval_device(pts, d_point_mean=mean.data_ptr(), d_point_m2=m2.data_ptr(), draws=(lo, hi), stream=st)
    shard_counts = [shard_range(forest.S, world, r)[1] - shard_range(forest.S, world, r)[0] for r in range(world)]
    counts, means, m2s = gather_moments(mean, m2, hi - lo, group, counts=shard_counts)
    var = torch.empty(pts.N, dtype=torch.float64, device=dev)
    w_ptr, w_len = (0, 0) if weights is None else (weights.data_ptr(), weights.numel())
    forest.merge_moments_device(counts, means.data_ptr(), m2s.data_ptr(), pts.N, w_ptr, w_len, 0, var.data_ptr(),
                                scaled=scaled, stream=st)
    return var
