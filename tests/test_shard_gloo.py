"""N>1 host logic on CPU: world_size-2 (and 3, ragged) gloo process groups exercise the image sharding
and the path's only collective (all-gather of detection records)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minddet_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_images, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(num_images * 100 * 5, dtype=torch.float32).reshape(num_images, 100, 5)
        a, b = shard.image_shard(num_images, world, rank)
        got = shard.gather_detections(full[a:b].clone(), num_images)
        got2 = shard.gather_detections(full[a:b].clone(), num_images, async_op=True).result()
        ok = got.shape == full.shape and torch.equal(got, full) and torch.equal(got2, full)
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok), a, b]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_images", [(2, 64), (2, 7), (3, 8)])
def test_gather_detections_gloo(tmp_path, world, num_images):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, num_images, str(tmp_path)), nprocs=world, join=True)
    covered = []
    for r in range(world):
        ok, a, b = np.load(tmp_path / f"ok{r}.npy")
        assert ok == 1
        covered += list(range(a, b))
    assert covered == list(range(num_images))          # shards are disjoint, ordered and cover the batch


def test_image_shard_properties():
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard.image_shard(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_sizes(n, w)
    with pytest.raises(ValueError):
        shard.image_shard(8, 2, 2)


def test_gather_without_process_group_is_identity():
    x = torch.rand(3, 100, 5)
    assert shard.gather_detections(x) is x
