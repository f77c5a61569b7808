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


def gather_detections(local, num_images=None, group=None):
    """All-gather per-image detection records.  local: (B_local, K, D) on this rank (B_local as given by
    `image_shard`).  Returns (num_images, K, D) in global image order on every rank.  One collective;
    ragged shards are padded to the largest shard so that the NCCL call stays a single fixed-size all-gather."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
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
    dist.all_gather_into_tensor(out, send, group=group)
    if all(s == bmax for s in sizes):
        return out
    return torch.cat([out[r * bmax:r * bmax + sizes[r]] for r in range(world)])
