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
PAIR_KEYS = ('f1', 'f2', 't12', 't21', 'ts12', 'ts21', 'm1', 'm2', 'g1', 'g2', 'h1', 'h2', 'kp1', 'kp2', 'p3d1', 'p3d2', 'dep1', 'dep2',
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


class BucketedGradAllReduce:
    """Gradient all-reduce (mean) of the trainable parameters, bucketed and overlapped with the backward pass.

    The reference trains under Lightning DDP (src/main.py:147-159): DDP packs gradients into ~25 MB buckets in
    reverse parameter order and all-reduces every bucket as soon as its last gradient has been produced, so the
    collective of the late layers runs while autograd is still working on the early ones.  This is the same
    protocol without the DDP wrapper (the loss ops of this package produce their gradients outside autograd's graph
    of the ViT, and ``find_unused_parameters`` is not needed): a post-accumulate-grad hook per parameter copies the
    gradient into its flat bucket; when a bucket is complete it is all-reduced on a side stream (NCCL) while the
    backward continues on the main stream.  ``finish()`` waits for the outstanding buckets, flushes buckets that
    stayed incomplete (parameters without a gradient this step count as zeros, like DDP's unused-parameter
    handling), divides by the world size, copies the means back into ``p.grad`` and optionally clips the global norm
    (Lightning ``gradient_clip_val=1.0``, src/main.py:158).

    Bucket size: DDP's 25 MB default is sized for PCIe / InfiniBand rings.  Over NVLink 5 / NVSwitch a 1.2 GB all-reduce
    takes 2 ms, so what shows up in the step is the fixed cost per collective (launch, stream hand-over, SMs taken
    from the backward kernels): measured on 2 x B200 with a ViT-L/14 full fine-tune step (tools/cfg5_probe.py), the
    exposed communication is 8.4 ms with 49 buckets of 25 MB, 3.5 ms with 12 x 100 MB and 0.76 ms with 3 x 400 MB.
    The default is therefore 256 MB.

    Works with any backend; with ``gloo`` (CPU tests) everything runs synchronously on the host.
    """

    def __init__(self, params, bucket_bytes=256 * 1024 * 1024, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.buckets = []          # [(flat tensor, [(param, offset, numel)])]
        self.where = {}            # param -> bucket index
        cur, cur_bytes = [], 0
        for p in reversed(self.params):        # gradients arrive roughly in reverse registration order
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self.pending = [0] * len(self.buckets)
        self.works = []
        self.launched = [False] * len(self.buckets)
        self.cuda = bool(self.params) and self.params[0].is_cuda
        self.stream = torch.cuda.Stream(device=self.params[0].device) if self.cuda else None
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.payload_bytes = sum(b.numel() * b.element_size() for b, _ in self.buckets)
        self.reset()

    def _close(self, plist):
        flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
        slots, off = [], 0
        for p in plist:
            slots.append((p, off, p.numel()))
            self.where[p] = len(self.buckets)
            off += p.numel()
        self.buckets.append((flat, slots))

    def reset(self):
        """Call before every backward pass."""
        self.pending = [len(slots) for _, slots in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.works = []
        self.filled = [set() for _ in self.buckets]

    def _launch(self, i):
        flat, _ = self.buckets[i]
        self.launched[i] = True
        if self.world == 1:
            return
        if self.cuda:
            # the bucket was filled on the current (backward) stream; the collective runs on the side stream
            self.stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self.stream):
                self.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _hook(self, p):
        i = self.where[p]
        flat, slots = self.buckets[i]
        for q, off, n in slots:
            if q is p:
                flat[off:off + n].copy_(p.grad.reshape(-1))
                break
        self.filled[i].add(id(p))
        self.pending[i] -= 1
        if self.pending[i] == 0:
            self._launch(i)

    def finish(self, clip_norm=None):
        """Wait for all buckets, write the mean gradients back into ``p.grad``; returns the global grad norm
        (a tensor) when ``clip_norm`` is given."""
        for i, (flat, slots) in enumerate(self.buckets):
            if not self.launched[i]:
                for q, off, n in slots:       # parameters that received no gradient this step contribute zeros
                    if id(q) not in self.filled[i]:
                        flat[off:off + n].zero_()
                self._launch(i)
        for w in self.works:
            w.wait()
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.params[0].device).wait_stream(self.stream)
        inv = 1.0 / self.world
        sq = None
        for flat, slots in self.buckets:
            if self.world > 1:
                flat.mul_(inv)
            if clip_norm is not None:
                s = flat.float().pow(2).sum()
                sq = s if sq is None else sq + s
        norm = None
        scale = None
        if clip_norm is not None and sq is not None:
            norm = sq.sqrt()
            scale = (clip_norm / (norm + 1e-6)).clamp(max=1.0)      # torch.nn.utils.clip_grad_norm_ semantics
        for flat, slots in self.buckets:
            if scale is not None:
                flat.mul_(scale.to(flat.dtype))
            for q, off, n in slots:
                if q.grad is None:
                    q.grad = flat[off:off + n].view_as(q).clone()
                else:
                    q.grad.copy_(flat[off:off + n].view_as(q))
        return norm

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []
