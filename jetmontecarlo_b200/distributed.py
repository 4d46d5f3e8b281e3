"""Multi-GPU plumbing: one process per GPU, samples sharded by global index,
histograms combined by ONE all-reduce (NCCL over NVLink on the GPU box; the
same code runs on gloo in the CPU tests).

Because sample i is a pure function of (seed, i) -- a counter RNG -- the union
of the ranks' samples is identical for every world size; the result differs
only by floating-point summation order.
"""
import torch
import torch.distributed as dist


def is_distributed(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def shard_bounds(n, rank, world):
    """[first, first + count) of rank ``rank``: contiguous, sizes differ by <= 1."""
    base, rem = divmod(int(n), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard(n, group=None):
    if not is_distributed(group):
        return 0, int(n)
    return shard_bounds(n, dist.get_rank(group), dist.get_world_size(group))


def pack_histograms(sum_w, sum_w2, count):
    """One float64 buffer [3, nb]; counts are exact in float64 up to 2^53."""
    return torch.stack([sum_w, sum_w2, count.to(torch.float64)])


def unpack_histograms(buf, sum_w, sum_w2, count):
    sum_w.copy_(buf[0])
    sum_w2.copy_(buf[1])
    count.copy_(buf[2].round().to(count.dtype))


def allreduce_histograms(sum_w, sum_w2, count, group=None):
    """SUM over ranks of (sum w, sum w^2, count), in place, one collective."""
    if not is_distributed(group):
        return
    buf = pack_histograms(sum_w, sum_w2, count)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    unpack_histograms(buf, sum_w, sum_w2, count)


def allreduce_minmax(minmax, group=None):
    """minmax[..., 0] = min, minmax[..., 1] = max (one pair, or one pair per parameter point of a grid), merged
    in place by ONE MAX all-reduce of [-min, max, has-NaN] rows; a NaN on any rank makes that pair NaN, as
    np.min / np.max of the whole sample set would be."""
    if not is_distributed(group):
        return
    inf = float("inf")
    mm = minmax.reshape(-1, 2)
    buf = torch.stack([-mm[:, 0], mm[:, 1]], dim=1)
    has_nan = torch.isnan(buf).any(dim=1, keepdim=True).to(buf.dtype)
    buf = torch.cat([torch.nan_to_num(buf, nan=-inf, posinf=inf, neginf=-inf), has_nan], dim=1).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
    out = torch.stack([-buf[:, 0], buf[:, 1]], dim=1)
    out[buf[:, 2] > 0] = float("nan")
    minmax.copy_(out.reshape(minmax.shape))


def gather_rows(rows, n_total, world, group=None):
    """Concatenate the ranks' row blocks (sizes from ``shard_bounds``: they differ by at most one row) into the full
    [n_total, ...] array on every rank: ONE all-gather of equally padded blocks, no reduction."""
    per = -(-int(n_total) // int(world))
    pad = torch.zeros((per,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    out = torch.empty((world * per,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    pieces = []
    for r in range(world):
        _, count = shard_bounds(n_total, r, world)
        pieces.append(out[r * per:r * per + count])
    return torch.cat(pieces, dim=0)
