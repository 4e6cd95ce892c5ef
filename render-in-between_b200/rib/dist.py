"""Multi-GPU plumbing: clips are independent, so they are sharded across ranks (one process per
GPU) with no collective in the forward; the only exchange is the final gather of the uint8 frames
onto one rank.  Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PendingGather:
    """Handle of an asynchronous gather_frames: `wait()` blocks the current stream (CUDA) or the host (CPU) until the
    frames have arrived and returns the gathered tensor on `dst` (None elsewhere)."""

    def __init__(self, works, out, keep):
        self._works, self._out, self._keep = works, out, keep

    def wait(self):
        for w in self._works:
            w.wait()
        self._works, self._keep = [], None
        return self._out


def gather_frames(local_u8, n_total, dst=0, group=None, out=None, async_op=False):
    """Gathers per-rank frame tensors [n_local, ...] (contiguous shards in rank order, see shard_range) onto `dst`.

    A true gather: only `dst` receives.  Every other rank sends its shard once, straight from `local_u8`; `dst`
    receives each shard into its slice of one [n_total, ...] tensor (`out`, allocated if None) and copies its own shard
    there, so ragged shards need no padding and nothing is concatenated afterwards.  With NCCL the sends / receives are
    one grouped P2P operation on NCCL's own stream: with async_op=True the calling stream is not blocked and the
    transfer overlaps whatever is launched next; call .wait() on the returned handle before reading `out`
    (and before `local_u8` is overwritten).

    Returns the [n_total, ...] tensor on `dst` and None elsewhere (or a PendingGather when async_op)."""
    if not (dist.is_available() and dist.is_initialized()):
        return PendingGather([], local_u8, None) if async_op else local_u8
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_total, rank, world)
    if local_u8.shape[0] != hi - lo:
        raise ValueError('rank %d holds %d items, its shard of %d is [%d, %d)' % (rank, local_u8.shape[0], n_total, lo, hi))
    local_u8 = local_u8.contiguous()
    ops, works = [], []
    if rank == dst:
        if out is None:
            out = torch.empty((n_total,) + tuple(local_u8.shape[1:]), dtype=local_u8.dtype, device=local_u8.device)
        elif tuple(out.shape) != (n_total,) + tuple(local_u8.shape[1:]) or not out.is_contiguous():
            raise ValueError('gather_frames: bad out tensor')
        for r in range(world):
            rlo, rhi = shard_range(n_total, r, world)
            if r == rank or rhi == rlo:
                continue
            ops.append(dist.P2POp(dist.irecv, out[rlo:rhi], r, group))
        out[lo:hi].copy_(local_u8, non_blocking=True)
    else:
        out = None
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, local_u8, dst, group))
    if ops:
        works = dist.batch_isend_irecv(ops)
    pending = PendingGather(works, out, local_u8)
    return pending if async_op else pending.wait()
