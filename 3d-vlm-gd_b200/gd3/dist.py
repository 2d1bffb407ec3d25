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


# per-pair entries of a batch dict (leading axis = pair; h1 / h2 / g1 / g2 may carry a layer axis in front of it);
# anything else (head parameters, a depth map shared by all pairs, scalars) is replicated
PAIR_KEYS = ('f1', 'f2', 't12', 't21', 'm1', 'm2', 'g1', 'g2', 'h1', 'h2', 'kp1', 'kp2', 'p3d1', 'p3d2', 'dep1', 'dep2',
             'depth_map1', 'depth_map2')


def shard_batch(batch, rank, world_size, pair_keys=PAIR_KEYS):
    """Slice the per-pair tensors of a batch dict (``pair_keys``) to this rank's chunk; everything else is replicated.

    Sharding goes by key, not by shape: a shared (H, W) depth map whose height happens to equal the number of pairs
    stays whole (it is 2-D), a per-pair (P, H, W) stack is sliced."""
    P = batch['f1'].shape[0]
    b, e = shard_range(P, rank, world_size)
    out = {}
    for k, v in batch.items():
        if k in pair_keys and torch.is_tensor(v):
            if k in ('depth_map1', 'depth_map2') and v.dim() == 2:
                out[k] = v
            elif k in ('g1', 'g2', 'h1', 'h2') and v.dim() == 4:
                out[k] = v[:, b:e]
            else:
                if v.shape[0] != P:
                    raise ValueError(f'shard_batch: {k} has leading size {v.shape[0]}, expected {P} pairs')
                out[k] = v[b:e]
        else:
            out[k] = v
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


def allreduce_mean_(tensor, group=None, local_pairs=None, num_pairs=None):
    """In-place mean over ranks (the DDP gradient semantics for the replicated depth-head parameters).

    With uneven shards pass ``local_pairs`` / ``num_pairs``: ``distillation_step`` scales a rank's gradient by
    1 / P_local, so the plain mean over ranks would over-weight the pairs of the smaller shards; the weighted form
    ``sum_r (P_r / P) g_r`` is the gradient of the mean loss over all P pairs."""
    if dist.is_available() and dist.is_initialized():
        if local_pairs is not None:
            tensor *= float(local_pairs) / float(num_pairs)
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
            tensor /= dist.get_world_size(group)
    return tensor
