"""Image sharding of the region path across the GPUs of one box.

Every stage of the path is per image, so a batch is cut into contiguous blocks of images, one block per
rank, and NOTHING crosses GPUs on the hot path.  The only collective is an all-gather of the final
fixed-size detection records (inference); training needs none here (the gradient all-reduce belongs to
the trainer: reference `nn.DistributedGradReducer`, pointpillars/src/pointpillars.py:900-910, dataset
sharding `de.DistributedSampler(device_num, rank)`, pointpillars/train.py:96).
"""
import torch
import torch.distributed as dist


def image_shard(num_images, world_size, rank):
    """Contiguous [start, stop) block of images owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(num_images, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(num_images, world_size):
    return [image_shard(num_images, world_size, r)[1] - image_shard(num_images, world_size, r)[0] for r in range(world_size)]


class PendingGather:
    """Handle of an asynchronous `gather_detections`: `.result()` waits (stream-side) and returns the gathered tensor."""

    def __init__(self, work, out, sizes, bmax):
        self.work, self.out, self.sizes, self.bmax = work, out, sizes, bmax

    def result(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        if all(s == self.bmax for s in self.sizes):
            return self.out
        return torch.cat([self.out[r * self.bmax:r * self.bmax + self.sizes[r]] for r in range(len(self.sizes))])


def gather_detections(local, num_images=None, group=None, async_op=False):
    """All-gather per-image detection records.  local: (B_local, K, D) on this rank (B_local as given by
    `image_shard`).  Returns (num_images, K, D) in global image order on every rank.  One collective;
    ragged shards are padded to the largest shard so that the NCCL call stays a single fixed-size all-gather.
    async_op=True returns a `PendingGather`: the collective then runs on NCCL's stream beside whatever the caller
    enqueues next (the detections of step i travel while step i+1 computes); `local` must not be overwritten
    before `.result()`."""
    if not (dist.is_available() and dist.is_initialized()):
        return PendingGather(None, local, [local.shape[0]], local.shape[0]) if async_op else local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if num_images is None:
        num_images = local.shape[0] * world
    sizes = shard_sizes(num_images, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} images, expected {sizes[rank]}")
    bmax = max(sizes)
    send = local.contiguous()
    if send.shape[0] < bmax:
        pad = torch.zeros((bmax - send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        send = torch.cat([send, pad])
    out = torch.empty((world * bmax,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    if async_op:
        return PendingGather(dist.all_gather_into_tensor(out, send, group=group, async_op=True), out, sizes, bmax)
    dist.all_gather_into_tensor(out, send, group=group)
    if all(s == bmax for s in sizes):
        return out
    return torch.cat([out[r * bmax:r * bmax + sizes[r]] for r in range(world)])
