"""Batch sharding across the GPUs of one node: one process + one engine per GPU, contiguous shards, no
data-path collective (every image is independent end to end -- net/v3.py:142 of the reference loops per
image and batch-norm runs in inference mode).  torch.distributed is only the control plane: a barrier,
the max-reduction of timings and the in-order gather of per-image results."""


def shard_bounds(n_items, world_size, rank):
    """Contiguous [start, stop) of `rank`; the first n_items % world_size ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_order(local_results):
    """Every rank passes the list of per-image results of its shard; every rank gets the full list in input order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_results)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, list(local_results))
    out = []
    for p in parts:
        out.extend(p)
    return out


def detect_sharded(detect_fn, images, world_size=None, rank=None):
    """Runs `detect_fn(images[start:stop])` on this rank's shard and returns all results in input order."""
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    start, stop = shard_bounds(len(images), world_size, rank)
    local = detect_fn(images[start:stop]) if stop > start else []
    return gather_in_order(local)
