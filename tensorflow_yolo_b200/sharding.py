"""Batch sharding across the GPUs of one node: one engine per GPU, contiguous shards, no data-path collective (every
image is independent end to end -- net/v3.py:142 of the reference loops per image and batch-norm runs in inference
mode).

Two drivers share `shard_bounds`:
  * in-process (the product: Yolo.test / detect_batch, net/yolo.py): `DevicePool`, one host thread per GPU (ctypes drops
    the GIL for the duration of every C-ABI call), results concatenated in input order;
  * one process per GPU (bench.py under torchrun): torch.distributed is only the control plane -- a barrier, the
    max-reduction of timings and the in-order gather of per-image results."""
import concurrent.futures
import os


def visible_devices():
    """Device ordinals the product may use: YB_DEVICES="0,2,3" | "all" (default: every visible device); the older
    YB_DEVICE=<n> still pins a single device."""
    from . import _lib
    spec = os.environ.get("YB_DEVICES")
    if spec is None and os.environ.get("YB_DEVICE") is not None:
        return [int(os.environ["YB_DEVICE"])]
    count = _lib.device_count()
    if spec is None or spec.strip().lower() in ("", "all"):
        return list(range(max(count, 1)))        # count == 0: device 0 is kept so that the engine reports the CUDA error
    devices = [int(v) for v in spec.split(",") if v.strip() != ""]
    if not devices:
        raise ValueError("YB_DEVICES={!r} names no device".format(spec))
    return devices


class DevicePool(object):
    """One worker thread per GPU.  `run(fn, items)` cuts `items` into contiguous shards (shard_bounds), calls
    fn(slot, shard) for slot i on thread i and returns the concatenated per-item results in input order."""

    def __init__(self, n_slots):
        self.n_slots = int(n_slots)
        self._workers = [concurrent.futures.ThreadPoolExecutor(max_workers=1) for _ in range(self.n_slots)]

    def each(self, fn):
        """fn(slot) on every worker thread, concurrently; returns the results in slot order."""
        futures = [w.submit(fn, i) for i, w in enumerate(self._workers)]
        return [f.result() for f in futures]

    def run(self, fn, items):
        n = len(items)
        used = min(self.n_slots, n)
        futures = []
        for slot in range(used):
            start, stop = shard_bounds(n, used, slot)
            futures.append(self._workers[slot].submit(fn, slot, items[start:stop]))
        out = []
        for f in futures:
            out.extend(f.result())
        return out

    def close(self):
        for w in self._workers:
            w.shutdown(wait=True)


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
