"""Multi-GPU plumbing: one process per GPU, the batch is sharded across ranks and no collective sits on the
data path (images are independent: no BatchNorm, nothing couples batch elements in G or CEM; SURVEY.md
§8e).  Replaces the thread-per-GPU ``nn.DataParallel`` scatter/gather of codes/models/networks.py:99-101.
``torch.distributed`` is used only for the optional gather of the 3-channel outputs."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of the n batch items owned by `rank` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def run_sharded(fn, batch, gather=True):
    """Applies fn to this rank's shard of `batch` (dim 0); optionally all-gathers the results in rank order."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(batch)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(batch.size(0), rank, world)
    out = fn(batch[lo:hi].contiguous()) if hi > lo else None
    if not gather:
        return out
    sizes = [shard_range(batch.size(0), r, world) for r in range(world)]
    shape = list(out.shape[1:]) if out is not None else None
    shapes = [None] * world
    dist.all_gather_object(shapes, shape)
    shape = next(s for s in shapes if s is not None)
    most = max(h - l for l, h in sizes)                      # all_gather needs equal sizes: pad, then trim
    mine = batch.new_zeros([most] + shape)
    if out is not None:
        mine[:out.size(0)] = out
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)], 0)
