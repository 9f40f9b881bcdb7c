"""The path's only cross-GPU exchange: observer statistics and scale gradients, combined with
torch.distributed collectives (NCCL over NVLink on GPUs, gloo in the CPU tests).

Activations shard by batch and weights are replicated, so the fake-quant kernels themselves need no
collective.  What must agree across ranks is O(channels) data:
  * observer statistics  - [C,4] = (min, max, max|x|, sum|x|): one all-gather of the blocks, folded locally,
  * sweep partial sums   - 80 squared-error sums per tensor: SUM,
  * scale gradients      - one flat buffer per step: SUM (DDP averages parameter grads itself),
  * per-channel WEIGHT observers (replicated weights, e.g. the 80-candidate sweep): each rank computes the
    qparams of a contiguous block of ceil(C / world) output channels and the blocks are all-gathered.
At world size 1 every function is the identity, i.e. bit-identical to the reference."""
import torch
import torch.distributed as dist

__all__ = ["world_size", "sync_stats", "sync_sse", "allreduce_grads_", "row_block", "rows_sharded"]

_enabled = True


def set_enabled(flag):
    """Observers all-reduce their statistics only while enabled (default on)."""
    global _enabled
    _enabled = bool(flag)


def world_size(group=None):
    if not _enabled or not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def sync_stats(stats, group=None):
    """stats [C,4] = (min, max, absmax, abssum) -> the statistics of the union of all ranks' tensors.
    ONE collective and no host synchronisation: the [C,4] blocks are all-gathered and folded locally (amin / amax /
    sum over the rank axis, in rank order - deterministic).  NaN statistics (NaN inputs) stay NaN: torch's amin / amax
    / sum propagate NaN, and a rank that saw a NaN reports NaN in all four columns (the statistics kernels do)."""
    w = world_size(group)
    if w == 1:
        return stats
    mine = stats.contiguous()
    parts = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(parts, mine, group=group)
    allr = torch.stack(parts, dim=0)                              # [world, C, 4]
    sm = allr[..., 3].sum(dim=0)
    out = torch.stack([allr[..., 0].amin(dim=0), allr[..., 1].amax(dim=0), allr[..., 2].amax(dim=0), sm], dim=1)
    # a NaN anywhere in a channel poisons its min / max / absmax too (what one pass over the union would give)
    return torch.where(torch.isnan(sm).unsqueeze(1), torch.full_like(out, float("nan")), out)


def sync_sse(sse, rows, group=None):
    """Sweep squared-error sums and the row count of l2_loss's mean, summed over ranks."""
    w = world_size(group)
    if w == 1:
        return sse, rows
    sse = sse.clone()
    dist.all_reduce(sse, op=dist.ReduceOp.SUM, group=group)
    return sse, rows * w


def allreduce_grads_(flat, average=False, group=None):
    """In-place SUM (or mean) of a flat scale-gradient buffer across ranks."""
    w = world_size(group)
    if w == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(w)
    return flat


def row_block(channels, rank, world):
    """Contiguous block of output channels owned by `rank`: [start, stop) with ceil(C / world) rows per rank."""
    per = (channels + world - 1) // world
    start = min(rank * per, channels)
    return start, min(start + per, channels)


def rows_sharded(rows2d, fn, group=None, min_rows_per_rank=2048):
    """Per-channel observer over REPLICATED rows, sharded by output channel: every rank runs
    `fn(block) -> tuple of [rows_in_block] tensors` on its block of rows and the results are all-gathered, so
    each row is computed once per job instead of once per rank.  Rows are independent and the weights are
    identical on all ranks, hence the gathered vectors equal fn(rows2d) bit for bit (fn must not let its
    launch geometry depend on the block size: F.sweep_channel takes the full row count for that).  Falls back
    to the local computation at world size 1 or when the matrix is too small to be worth an exchange:
    measured on 4 B200s (profiles/r01_sharded_observer_n4.json), sharding each of ResNet-50's 54 weight
    tensors (<= 2048 rows) costs 108 all-gathers and takes 5.97 ms against 3.51 ms for sweeping everything
    on every rank - the sweep of a CNN layer is launch-latency-bound - hence the high default threshold."""
    w = world_size(group)
    channels = rows2d.shape[0]
    if w == 1 or channels < w * min_rows_per_rank:
        return fn(rows2d)
    rank = dist.get_rank(group)
    per = (channels + w - 1) // w
    start, stop = row_block(channels, rank, w)
    mine = fn(rows2d[start:stop]) if stop > start else None
    out = []
    n_out = len(mine) if mine is not None else None
    if n_out is None:                       # an empty block still has to take part in the collectives
        probe = fn(rows2d[:1])
        n_out = len(probe)
        mine = tuple(p[:0] for p in probe)
    for k in range(n_out):
        pad = torch.zeros(per, dtype=mine[k].dtype, device=rows2d.device)
        pad[:stop - start] = mine[k].reshape(-1)
        gathered = torch.empty(per * w, dtype=pad.dtype, device=rows2d.device)
        if rows2d.is_cuda:
            dist.all_gather_into_tensor(gathered, pad, group=group)
        else:                               # gloo (CPU tests) has no all_gather_into_tensor
            dist.all_gather(list(gathered.split(per)), pad, group=group)
        out.append(gathered[:channels])
    return tuple(out)
