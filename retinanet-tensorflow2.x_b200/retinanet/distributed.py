"""Image sharding for batched / multi-GPU inference (reference: notebooks/multi_gpu_inference.ipynb cell 6 —
MirroredStrategy, one shard of images per replica, results gathered to the host; executor.py:397-398
strategy.gather(axis=0)).

One process per GPU.  Images are independent on this path, so each rank post-processes its own shard and there is no
collective on the hot path; `gather_detections` is the optional off-path gather of the four small output tensors
(about 2.4 KB per image) with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(num_images, rank, world_size):
    """Contiguous shard [lo, hi) of rank `rank`: the first (num_images % world_size) ranks take one extra image."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError('bad rank/world_size: {}/{}'.format(rank, world_size))
    base, extra = divmod(int(num_images), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(predictions, rank, world_size):
    """Slices every tensor of a (possibly nested) prediction dict along the batch axis."""
    def cut(t):
        lo, hi = shard_range(t.shape[0], rank, world_size)
        return t[lo:hi]
    return {k: (shard_batch(v, rank, world_size) if isinstance(v, dict) else cut(v)) for k, v in predictions.items()}


def gather_detections(detections, num_images=None, group=None):
    """All-gathers {'boxes','scores','classes','valid_detections'} of every rank, concatenated in rank order along
    axis 0 (= strategy.gather(axis=0)).  Shards may be ragged (num_images not divisible by world size)."""
    if not dist.is_available() or not dist.is_initialized():
        return detections
    world = dist.get_world_size(group)
    local = int(next(iter(detections.values())).shape[0])
    counts = [None] * world
    dist.all_gather_object(counts, local, group=group)
    most = max(counts)
    out = {}
    for key, t in detections.items():
        pad = torch.zeros((most,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:local] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[key] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    if num_images is not None and out['valid_detections'].shape[0] != num_images:
        raise RuntimeError('gathered {} images, expected {}'.format(out['valid_detections'].shape[0], num_images))
    return out
