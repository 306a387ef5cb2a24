"""CPU: the N>1 host logic (sharding, in-order gather, max-reduction of timings) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from tensorflow_yolo_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 128, 129, 512):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    images = np.arange(7 * 4, dtype=np.float32).reshape(7, 4)        # 7 "images": ragged over 2 ranks (4 + 3)

    def fake_detect(shard):                                           # stands in for Engine.forward + detect
        return [{"first": float(im[0]), "rank": rank} for im in shard]

    res = sharding.detect_sharded(fake_detect, images)
    slowest = sharding.max_over_ranks(10.0 + rank)
    dist.barrier()
    np.save(os.path.join(out_dir, "r%d.npy" % rank),
            np.asarray([[r["first"], r["rank"]] for r in res] + [[slowest, -1]]))
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b)                                       # every rank sees the same, ordered result
    assert a[:-1, 0].tolist() == [0.0, 4.0, 8.0, 12.0, 16.0, 20.0, 24.0]
    assert a[:-1, 1].tolist() == [0, 0, 0, 0, 1, 1, 1]                # contiguous shards: 4 + 3
    assert a[-1, 0] == 11.0                                           # max over ranks of (10 + rank)


def test_device_pool_keeps_input_order_and_balances():
    """The in-process driver (one host thread per GPU, sharding.DevicePool): contiguous shards, input order kept."""
    import threading
    pool = sharding.DevicePool(3)
    seen = {}

    def work(slot, shard):
        seen.setdefault(slot, []).append((threading.get_ident(), list(shard)))
        return [v * v for v in shard]
    for n in (0, 1, 2, 3, 7, 10):
        items = list(range(n))
        assert pool.run(work, items) == [v * v for v in items]
    sizes = [len(s) for _, s in (seen[0][-1], seen[1][-1], seen[2][-1])]
    assert sizes == [4, 3, 3]                                     # 10 items over 3 slots
    assert len({t for calls in seen.values() for t, _ in calls}) == 3        # one thread per slot, always the same
    assert pool.each(lambda slot: slot * 2) == [0, 2, 4]
    pool.close()


def test_visible_devices_parsing(monkeypatch):
    monkeypatch.setenv("YB_DEVICES", "2, 0")
    assert sharding.visible_devices() == [2, 0]
    monkeypatch.delenv("YB_DEVICES")
    monkeypatch.setenv("YB_DEVICE", "1")
    assert sharding.visible_devices() == [1]
    monkeypatch.delenv("YB_DEVICE")
    assert sharding.visible_devices() == [0]                      # no GPU here: device 0 stays so the engine reports the CUDA error
