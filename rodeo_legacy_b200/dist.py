"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed).

The path shards trivially: systems are independent, so rank r solves the contiguous slice
[lo, hi) of the batch with no communication, and ONE collective at the end gathers per-system
results (8 bytes per system for a log-likelihood sweep).  Nothing else crosses NVLink.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_sizes", "gather_batch", "sharded_loglik"]


def shard_sizes(batch, world_size):
    """Contiguous, balanced split: the first (batch % world_size) ranks get one extra system."""
    q, r = divmod(int(batch), int(world_size))
    return [q + (1 if i < r else 0) for i in range(world_size)]


def shard_range(batch, rank, world_size):
    sizes = shard_sizes(batch, world_size)
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def gather_batch(local, batch, group=None):
    """all_gather of a per-system tensor whose first dim is this rank's shard (ragged shards
    are padded to the largest one).  Returns the full (batch, ...) tensor on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(batch, world)
    width = max(sizes)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[i * width: i * width + s] for i, s in enumerate(sizes)], dim=0)


def sharded_loglik(local_loglik_fn, x0, theta, group=None):
    """Evaluate ``local_loglik_fn(x0_shard, theta_shard) -> (b,)`` on this rank's slice of the
    batch and gather the per-system values from all ranks."""
    batch = x0.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(batch, rank, world)
    local = local_loglik_fn(x0[lo:hi], None if theta is None else theta[lo:hi])
    return gather_batch(local, batch, group)
