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


def _on_device(*tensors):
    return all(t.is_cuda and t.is_contiguous() for t in tensors)


def _stream():
    import ctypes as C
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_histograms(sum_w, sum_w2, count):
    """One float64 buffer [3, nb]; counts are exact in float64 up to 2^53.  On the device: one kernel of the library
    (jmc_pack_histograms); host tensors (the gloo tests of this logic) are packed by torch."""
    if _on_device(sum_w, sum_w2, count) and sum_w.dtype == torch.float64 and count.dtype == torch.int64:
        from . import _lib
        buf = torch.empty((3,) + tuple(sum_w.shape), dtype=torch.float64, device=sum_w.device)
        _lib.check(_lib.lib.jmc_pack_histograms(sum_w.data_ptr(), sum_w2.data_ptr(), count.data_ptr(), sum_w.numel(),
                                                buf.data_ptr(), 0, _stream()))
        return buf
    return torch.stack([sum_w, sum_w2, count.to(torch.float64)])


def unpack_histograms(buf, sum_w, sum_w2, count):
    if _on_device(buf, sum_w, sum_w2, count) and sum_w.dtype == torch.float64 and count.dtype == torch.int64:
        from . import _lib
        _lib.check(_lib.lib.jmc_pack_histograms(sum_w.data_ptr(), sum_w2.data_ptr(), count.data_ptr(), sum_w.numel(),
                                                buf.data_ptr(), 1, _stream()))
        return
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
    if _on_device(minmax) and minmax.dtype == torch.float64:
        from . import _lib
        n_pairs = minmax.numel() // 2
        buf = torch.empty((n_pairs, 3), dtype=torch.float64, device=minmax.device)
        _lib.check(_lib.lib.jmc_pack_minmax(minmax.data_ptr(), n_pairs, buf.data_ptr(), 0, _stream()))
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
        _lib.check(_lib.lib.jmc_pack_minmax(minmax.data_ptr(), n_pairs, buf.data_ptr(), 1, _stream()))
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
    rows = rows.contiguous()
    pad = torch.empty((per,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]].copy_(rows)                 # contiguous blocks: device-to-device copies, no kernel
    # (a rank with one row less leaves its padding row as allocated: it is never read back)
    out = torch.empty((world * per,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    full = torch.empty((int(n_total),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    at = 0
    for r in range(world):
        _, count = shard_bounds(n_total, r, world)
        full[at:at + count].copy_(out[r * per:r * per + count])
        at += count
    return full
