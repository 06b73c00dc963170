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


def ordered_sum(rows: torch.Tensor, stream=None) -> torch.Tensor:
    """Sum over dim 0 in a fixed order (deterministic, identical on every rank).  CUDA tensors: ONE launch of the
    library's fixed-order row reduction (bartint_sum_rows_device) instead of world - 1 elementwise kernels."""
    if rows.is_cuda and rows.dtype == torch.float64 and rows.dim() == 2 and rows.is_contiguous():
        from . import _lib
        out = torch.empty(rows.shape[1], dtype=torch.float64, device=rows.device)
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        rc = _lib.load().bartint_sum_rows_device(rows.data_ptr(), rows.shape[0], rows.shape[1], out.data_ptr(), st)
        if rc != 0:
            raise RuntimeError(_lib.load().bartint_last_error().decode())
        return out
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
    """Draw-sharded exchange: (draw counts per rank, means [world, C], M2s [world, C]) -- one all-gather for both."""
    both = all_gather_rows(torch.stack([local_mean, local_m2]), group)          # [world, 2, C]
    means, m2s = both[:, 0].contiguous(), both[:, 1].contiguous()
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
def point_sharded_draw_means(forest, pts_local, group=None, scaled=False, stream=None, n_total=None, no_bitslice=False):
    """rowMeans(predict(model, X)) with X's rows sharded across ranks.  Returns (draw_mean[S], mean_var[2]) on device."""
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    sums = torch.empty(forest.S, dtype=torch.float64, device=dev)
    forest.eval_device(pts_local, d_draw_sum=sums.data_ptr(), stream=st, no_bitslice=no_bitslice)
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


# ---------------------------------------------------------------------------------------------
# draw-sharded EXPORT + replicated forest: every rank parses only its share of the posterior
# ---------------------------------------------------------------------------------------------
def replicate_forest(local_forest, group=None, stream=None):
    """All-gather the device images of the ranks' local sub-forests (bartint_forest_pack / _unpack): returns the list of
    `world` forests in rank order; entry `rank` is local_forest itself.  MB-sized payload over NVLink instead of every
    rank exporting the whole posterior on its share of the host cores."""
    from .forest import Forest
    rank, world = world_info(group)
    if world == 1:
        return [local_forest]
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    hb, db = local_forest.image_size()
    sizes = all_gather_rows(torch.tensor([hb, db], dtype=torch.int64, device=dev), group).cpu().numpy()
    max_hb, max_db = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    blob = torch.empty(max_db, dtype=torch.uint8, device=dev)
    header = local_forest.pack(blob.data_ptr(), db, stream=st)
    hdr = torch.zeros(max_hb, dtype=torch.uint8)
    hdr[:hb] = torch.from_numpy(header)
    all_hdr = all_gather_rows(hdr.to(dev), group).cpu().numpy()
    all_blob = all_gather_rows(blob, group)
    out = []
    for r in range(world):
        if r == rank:
            out.append(local_forest)
        else:
            out.append(Forest.unpack(all_hdr[r, :int(sizes[r, 0])], all_blob[r].data_ptr(), int(sizes[r, 1]), stream=st))
    return out


def sharded_design_step(tree_strings, tree_fits, x_train, cut_lists, y_min, y_max, mc_local, candidates, weights=None,
                        measure="uniform", group=None, n_total=None, n_draws_integral=None):
    """One sequential-design step on `world` ranks from the dbarts state every rank holds (the multi-GPU counterpart of
    design_step_dbarts):
      export      draw-sharded: rank r parses draws [lo_r, hi_r) only, then the finished sub-forests are all-gathered
      integrals   own draws, all-gathered                                       (BARTInt.R:259-262)
      candidates  own draws on all candidates, (n, mean, M2) all-gathered, Chan merge   (BARTInt.R:312-314)
      MC points   point-sharded: own rows (StagedMatrix) on ALL draws, per-draw sums all-gathered (bartMean.R:74-82)
    n_total: the global number of MC rows, if the caller knows it (saves one exchange + host wait).
    n_draws_integral: integrate only the first so many draws (the reference's quirk B1: min(ntree, S), BARTInt.R:216);
    None = all draws.  Mean and sd are then taken over those draws, like bartint_design_step_dbarts098.
    Returns a dict of host arrays like design_step_dbarts (identical on every rank)."""
    from .forest import Forest, WireStrings
    rank, world = world_info(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream
    wire = tree_strings if isinstance(tree_strings, WireStrings) else WireStrings(tree_strings)
    S, m = wire.S, wire.m
    x_train = np.asarray(x_train, np.float64)
    n = x_train.shape[0]
    fits = np.asarray(tree_fits, np.float64).reshape(n, m, S)
    bounds = [shard_range(S, world, r) for r in range(world)]
    lo, hi = bounds[rank]
    mc_local.flush()          # this rank's forest upload is ~1 ms away: let the whole matrix travel meanwhile
    f_loc = Forest.from_dbarts(wire.slice_draws(lo, hi), fits[:, :, lo:hi], x_train, cut_lists, y_min, y_max)
    # Exchange the sub-forest images FIRST, on an otherwise idle GPU: a collective queued behind the evaluation
    # kernel cannot start until that kernel's CTAs (resident for its whole duration) leave the SMs -- measured 4-9 ms
    # for the exchange when it was overlapped with the own-draw evaluation, 0.8 ms alone.
    forests = replicate_forest(f_loc, group, stream=st)
    # own draws: exact integrals + candidate moments (small)
    n_int = S if not n_draws_integral else max(1, min(int(n_draws_integral), S))
    integ = torch.zeros(S, dtype=torch.float64, device=dev)
    if min(hi, n_int) > lo:          # this rank's draws inside the integrated range
        f_loc.exact_integrals_device(integ[lo:hi].data_ptr(), measure, n_draws_used=min(hi, n_int) - lo, stream=st)
    C_ = candidates.N
    pc = candidates.points(f_loc, stream=st)
    mean = torch.zeros(C_, dtype=torch.float64, device=dev)
    m2 = torch.zeros(C_, dtype=torch.float64, device=dev)
    f_loc.eval_device(pc, d_point_mean=mean.data_ptr(), d_point_m2=m2.data_ptr(), stream=st)
    # MC points: own rows on every rank's draws (point-sharded)
    pm = mc_local.points(f_loc, stream=st)
    sums = torch.empty(S, dtype=torch.float64, device=dev)
    for r, (a, b) in enumerate(bounds):
        if b > a:
            forests[r].eval_device(pm, d_draw_sum=sums[a:b].data_ptr(), stream=st)
    if n_total is None:       # costs a device->host wait behind the evaluation; pass n_total when it is known
        n_total = int(all_gather_rows(torch.tensor([pm.N], dtype=torch.int64, device=dev), group).sum().item()) if world > 1 else pm.N
    total, _ = allreduce_draw_sums(sums, pm.N, group, n_total=n_total)
    draw_mean = torch.empty(S, dtype=torch.float64, device=dev)
    mv = torch.empty(2, dtype=torch.float64, device=dev)
    f_loc.finalize_draws_device(total.data_ptr(), S, n_total, draw_mean.data_ptr(), mv.data_ptr(), stream=st)
    # candidates: Chan merge of the ranks' moments in rank order
    counts, means, m2s = gather_moments(mean, m2, hi - lo, group, counts=[b - a for a, b in bounds])
    var = torch.empty(C_, dtype=torch.float64, device=dev)
    w = None if weights is None else torch.as_tensor(np.atleast_1d(np.asarray(weights, np.float64)), device=dev)
    f_loc.merge_moments_device(counts, means.data_ptr(), m2s.data_ptr(), C_, 0 if w is None else w.data_ptr(),
                               0 if w is None else w.numel(), 0, var.data_ptr(), stream=st)
    # integrals: each rank filled its own range; the rest is zero, so a rank-ordered sum of the gathered rows is a gather
    integ_all = ordered_sum(all_gather_rows(integ, group)) if world > 1 else integ
    integ_res = torch.empty(S, dtype=torch.float64, device=dev)
    integ_ms = torch.empty(2, dtype=torch.float64, device=dev)
    f_loc.finalize_integrals_device(integ_all.data_ptr(), n_int, integ_res.data_ptr(), integ_ms.data_ptr(), stream=st)
    host = torch.cat([integ_all, integ_ms, var, draw_mean, mv]).cpu().numpy()        # one device->host copy
    out = {"integrals": host[:n_int], "integral_mean_sd": host[S:S + 2], "cand_var": host[S + 2:S + 2 + C_],
           "draw_mean": host[S + 2 + C_:2 * S + 2 + C_], "draw_mean_var": host[2 * S + 2 + C_:]}
    pm.close(); pc.close()
    for r, f in enumerate(forests):
        f.close()
    return out
