"""Data-parallel sharding of image pairs.

Every loss is a function of one image pair, so the path shards into independent units: rank r owns the
contiguous chunk [r*P/G, (r+1)*P/G) of a global batch, one process per GPU, and there is NO collective
inside the loss path.  The reference does the same with Lightning DDP and one pair per rank
(src/main.py:147-151); its only collective is the gradient all-reduce of the trainable parameters after
the ViT backward.  ``gather_pair_losses`` / ``allreduce_mean_`` are the two thin torch.distributed
helpers a trainer needs around the loss op: per-pair loss values for logging, and the mean of the
(small) depth-head gradients across ranks.
"""
import torch
import torch.distributed as dist


def shard_range(num_pairs, rank, world_size):
    """Contiguous [begin, end) of the pairs owned by ``rank``; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank {rank} / world size {world_size}')
    base, rem = divmod(num_pairs, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world_size):
    """Slice every per-pair tensor of a batch dict to this rank's chunk (the head is replicated)."""
    P = batch['f1'].shape[0]
    b, e = shard_range(P, rank, world_size)
    out = {}
    for k, v in batch.items():
        out[k] = v[b:e] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == P else v
    return out


def gather_pair_losses(local_losses, num_pairs, group=None):
    """All-gather per-pair loss values (P_local,) into the global order (P,) on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_losses
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(num_pairs, r, world) for r in range(world)]
    width = max(e - b for b, e in sizes)
    pad = torch.zeros(width, dtype=local_losses.dtype, device=local_losses.device)
    pad[:local_losses.numel()] = local_losses
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    assert sizes[rank][1] - sizes[rank][0] == local_losses.numel()
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)])


def allreduce_mean_(tensor, group=None):
    """In-place mean over ranks (the DDP gradient semantics for the replicated depth-head parameters)."""
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
        tensor /= dist.get_world_size(group)
    return tensor
