"""Image-sharded inference across the GPUs of one box.

Each image's output depends only on that image (InstanceNorm is per-sample, networks.py:31), so
inference partitions by image with NO data-path collective: image i goes to rank i mod N, one process
per GPU, weights replicated. The only exchange is an optional gather of per-rank results/metrics at the
end, over torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch


def shard_indices(num_items, rank, world_size):
    """Indices of the items rank `rank` processes: i with i % world_size == rank (round robin)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, num_items, world_size))


def gather_results(local_values, num_items, rank, world_size, group=None):
    """All-gathers one float per locally processed item into a (num_items,) tensor in item order.

    local_values[j] belongs to item shard_indices(...)[j]. Works on any backend; with world_size == 1 no
    collective is issued.
    """
    idx = shard_indices(num_items, rank, world_size)
    if len(local_values) != len(idx):
        raise ValueError("rank %d produced %d values for %d items" % (rank, len(local_values), len(idx)))
    local = torch.as_tensor(local_values, dtype=torch.float64).flatten()
    out = torch.full((num_items,), float("nan"), dtype=torch.float64, device=local.device)
    if world_size == 1:
        out[idx] = local
        return out
    import torch.distributed as dist
    per_rank = (num_items + world_size - 1) // world_size
    padded = torch.full((per_rank,), float("nan"), dtype=torch.float64, device=local.device)
    padded[: len(idx)] = local
    bucket = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(bucket, padded, group=group)
    for r in range(world_size):
        ridx = shard_indices(num_items, r, world_size)
        out[ridx] = bucket[r][: len(ridx)]
    return out
