"""Multi-GPU plumbing: clips are independent, so they are sharded across ranks (one process per
GPU) with no collective in the forward; the only exchange is the final gather of the uint8 frames
(SURVEY.md §8e).  Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_frames(local_u8, n_total, dst=0, group=None):
    """Gathers per-rank uint8 frame tensors [n_local, ...] (contiguous shards, rank order) onto `dst`.

    Returns the concatenated [n_total, ...] tensor on `dst` and None elsewhere.  Shards may be ragged:
    they are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_u8
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((n_max,) + tuple(local_u8.shape[1:]), dtype=local_u8.dtype, device=local_u8.device)
    pad[:local_u8.shape[0]] = local_u8
    if dist.get_backend(group) == 'nccl':
        out = torch.empty((world * n_max,) + tuple(local_u8.shape[1:]), dtype=local_u8.dtype, device=local_u8.device)
        dist.all_gather_into_tensor(out, pad, group=group)       # NVLink/NVSwitch; frames are tiny next to compute
        parts = list(out.view((world, n_max) + tuple(local_u8.shape[1:])))
    else:
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
