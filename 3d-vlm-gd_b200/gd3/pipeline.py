"""One fused geometric-distillation step over a batch of image pairs (product path, no autograd).

Mirrors steps 4-7 of the reference's ``training_step`` (src/finetune_timm_mast3r.py:636-653, vggt
:614-625): depth losses (cross-view L1 + intra-view ranking), cost-volume KL, Smooth-AP matching and
their weighted sum, for P pairs at once.  Everything runs through the C ABI of lib3dgd.so; torch is
used for buffers and the stream only.  The total is the mean over pairs of

    w_ap * ap + w_depth * l1 + w_intra * rank + w_kl * kl

(the reference's per-rank loss with one pair per rank; its DDP then averages gradients over ranks).
"""
import torch

from . import _lib, ops

_F32 = torch.float32

# constructor defaults of the reference modules: mast3r (src/finetune_timm_mast3r.py:79-84) trains with
# depth weight 0, vggt (src/finetune_timm_vggt.py:86-89) with all weights 1
DEFAULT_WEIGHTS = {
    'mast3r': dict(ap=1.0, depth=0.0, intra=1.0, kl=1.0),
    'vggt': dict(ap=1.0, depth=1.0, intra=1.0, kl=1.0),
}


_SIDE_STREAMS = {}


def _side_streams(dev):
    """Three long-lived side streams per device for the independent loss branches of a step: KL, Smooth-AP, and a
    HIGH-priority one for the depth-ranking pipeline -- its pair kernel is 70 % of the step, so its CTAs should never
    queue behind the other branches' (3.24 -> 3.21 ms per step at cfg2)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev),
                              torch.cuda.Stream(device=dev, priority=-1)]
    return _SIDE_STREAMS[key]


def distillation_step(batch, variant='mast3r', grid=None, patch_size=14, backward=True, weights=None,
                      pairs_per_group=0, depth_threshold=0.05, thr_neg=0.1, temp=0.01, parallel_branches=False):
    """Run the three distillation losses (+ L1) forward and backward for a batch of pairs.

    batch (all CUDA tensors):
      f1, f2     (P, N, C)  student patch features for the cost volume (fp32 / bf16)
      t12, t21   (P, N, N)  teacher volumes, fp32 -- or fp16 from ``ops.pack_teacher`` together with its row statistics
                 ts12, ts21 (P, 3, N);  m1, m2 (P, N) bool patch masks
      g1, g2     (P, N, C) or (L, P, N, C)  token maps the matching descriptors are sampled from (L maps: their mean)
      h1, h2     optional, same shapes: token maps the depth-head features are sampled from (default: g1, g2)
      kp1, kp2   (P, K, 2)  pixel keypoints;  p3d1, p3d2 (P, K, 3);  dep1, dep2 (P, K) keypoint depths
      Derived on the device when absent (one ``gd3_kp_prepare`` launch per view): m1 / m2 = patches holding a keypoint
      (src/finetune_timm_mast3r.py:516-519), dep1 / dep2 = 3 x 3 window depths at the keypoints from ``depth_map1`` /
      ``depth_map2`` ((P, H, W) or one shared (H, W) map; src/finetune_timm_mast3r.py:482-483).
      head       dict W1, b1, gamma, beta, w2, b2 (+ use_tanh, ln_eps) of the depth-difference head
    parallel_branches: the three losses are independent until the final scatter; with True the KL, Smooth-AP and
      depth-ranking pipelines are enqueued on three side streams (the ranking one with high priority; under CUDA-graph
      capture they become parallel branches of the graph).  With the round-2 pair kernel (2 CTAs x 4 warps
      per SM: 57 K registers, 136 KB of shared memory) the other pipelines find room next to it and in its tail: 3.24
      instead of 3.32 ms per step at cfg2; the earlier pair kernel owned every register of the SM and the
      branches gained nothing (3.92 against 3.88 ms).  The default stays False for eager callers (the side streams cost
      event traffic per call); ``GraphedStep`` and the benchmark turn it on.
    Returns dict: kl, ap, rank, l1 (each (P,)), total (0-d) and, if backward, ``grads`` with f1, f2 (feature
    dtype), g1, g2 (fp32; h1, h2 too when given) and head (packed [W1 | b1 | gamma | beta | w2 | b2]).
    """
    w = dict(DEFAULT_WEIGHTS[variant])
    if weights:
        w.update(weights)
    f1, f2 = batch['f1'], batch['f2']
    P, N, C = f1.shape
    ph, pw = grid
    geom = (ph, pw, ph * patch_size, pw * patch_size, patch_size, patch_size)
    inv_p = 1.0 / max(P, 1)
    head = batch['head']
    params = (head['W1'], head['b1'], head['gamma'], head['beta'], head['w2'], head['b2'])
    dev = f1.device
    out = {}

    # ---- keypoint descriptors / features from the token maps (K3) ----
    # g1 / g2: maps the matching descriptors are sampled from (the reference's refine_conv output, normalised);
    # h1 / h2 (optional, default g1 / g2): maps the depth-head features are sampled from (the reference's mean of
    # blocks 4..7, src/finetune_timm_mast3r.py:271-277 -- pass the (L, P, N, C) stack, the mean is folded into the sample)
    g1, g2 = batch['g1'], batch['g2']
    h1, h2 = batch.get('h1', g1), batch.get('h2', g2)
    kp1 = batch['kp1'].to(_F32).contiguous()
    kp2 = batch['kp2'].to(_F32).contiguous()
    K = kp1.shape[1]
    lays = {}
    for name, t in (('g1', g1), ('g2', g2), ('h1', h1), ('h2', h2)):
        L_, P_, N_, C_, st, gst = ops.token_layout(t)
        if P_ != P or N_ != ph * pw:
            raise ValueError(f'distillation_step: {name} is {tuple(t.shape)}, expected (..., {P}, {ph * pw}, C) for a '
                             f'{ph} x {pw} patch grid')
        lays[name] = (L_, P_, N_, C_, st, gst)
    if lays['g1'][3] != lays['g2'][3] or lays['h1'][3] != lays['h2'][3]:
        raise ValueError('distillation_step: the two views must share the channel count of each token map')
    Cd, Ch = lays['g1'][3], lays['h1'][3]
    pstr = (2 * K * Ch, Ch, 1)
    # patch masks and keypoint depths the caller did not supply
    prepared = {}
    for v, kp in (('1', kp1), ('2', kp2)):
        need_m, need_d = ('m' + v) not in batch, ('dep' + v) not in batch
        if need_m or need_d:
            m, d = _lib.kp_prepare(kp, geom[2], geom[3], patch_size=patch_size if need_m else None,
                                   depth=batch['depth_map' + v] if need_d else None)
            prepared['m' + v], prepared['dep' + v] = m, d
    m1 = batch['m1'] if 'm1' in batch else prepared['m1']
    m2 = batch['m2'] if 'm2' in batch else prepared['m2']
    dep1 = batch['dep1'] if 'dep1' in batch else prepared['dep1']
    dep2 = batch['dep2'] if 'dep2' in batch else prepared['dep2']
    depths = torch.stack([dep1.to(_F32), dep2.to(_F32)], dim=1).reshape(2 * P, K)

    # teacher volumes: fp32, or the producers' packed form (fp16 volume t12 / t21 + row statistics ts12 / ts21)
    t12 = (batch['t12'], batch['ts12']) if 'ts12' in batch else batch['t12']
    t21 = (batch['t21'], batch['ts21']) if 'ts21' in batch else batch['t21']
    w_rank = _const_vector(2 * P, 0.5 * w['intra'] * inv_p, dev)
    w_l1 = _const_vector(P, w['depth'] * inv_p, dev)
    main = torch.cuda.current_stream(dev)
    if parallel_branches:
        s_kl, s_ap, s_rank = _side_streams(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        for side in (s_kl, s_ap, s_rank):
            side.wait_event(fork)
    else:
        s_kl = s_ap = s_rank = main
    # the zero-initialised token-gradient maps of the backward scatter: filled on the caller's stream, which has nothing
    # else to do until the branches join
    shared1, shared2 = h1 is g1, h2 is g2
    gmaps = {}
    if backward:
        for name, t, wanted in (('g1', g1, True), ('g2', g2, True), ('h1', h1, not shared1), ('h2', h2, not shared2)):
            if wanted:
                L_, P_, N_, C_, _, _ = lays[name]
                gmaps[name] = torch.zeros((L_, P_, N_, C_) if t.dim() == 4 else (P_, N_, C_), dtype=_F32, device=dev)

    # ---- dense cost-volume KL (K1), side stream ----
    with torch.cuda.stream(s_kl):
        kl, gf1, gf2 = ops.cost_kl_raw(f1, f2, t12, t21, m1, m2, variant, grad_scale=w['kl'] * inv_p,
                                       want_grad=backward, pairs_per_group=pairs_per_group)
    # ---- Smooth-AP (K2), side stream: samples its own (normalised) descriptors ----
    with torch.cuda.stream(s_ap):
        d1, inv1, ostr = ops.sample_fwd_raw(g1, lays['g1'][:5], geom, kp1, True)
        d2, inv2, _ = ops.sample_fwd_raw(g2, lays['g2'][:5], geom, kp2, True)
        ap, gd1, gd2 = ops.smooth_ap_raw(d1, d2, batch['p3d1'], batch['p3d2'], variant, temp, thr_neg,
                                         grad_scale=w['ap'] * inv_p, want_grad=backward)
    # ---- relative depth: ranking on both views + cross-view L1 (K4): the caller's stream, or the high-priority side stream ----
    with torch.cuda.stream(s_rank):
        # depth features of both views interleaved as sets (2p, 2p+1) = (view 1, view 2) of pair p
        kf = torch.empty(P, 2, K, Ch, dtype=_F32, device=dev)
        ops.sample_fwd_raw(h1, lays['h1'][:5], geom, kp1, False, out=kf[:, 0], out_strides=pstr)
        ops.sample_fwd_raw(h2, lays['h2'][:5], geom, kp2, False, out=kf[:, 1], out_strides=pstr)
        lr, l1, gkf, gparams = ops.depth_head_raw(kf.reshape(2 * P, K, Ch), depths, params,
                                                  head.get('use_tanh', True), head.get('ln_eps', 1e-5), 0,
                                                  depth_threshold, 0.05, False, w_rank, w_l1, backward)
    if parallel_branches:
        for side, outs in ((s_kl, (kl, gf1, gf2)), (s_ap, (ap, gd1, gd2, d1, d2, inv1, inv2)),
                           (s_rank, (lr, l1, gkf, gparams, kf))):
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
            if not torch.cuda.is_current_stream_capturing():
                for t in outs:       # produced on a side stream, consumed (and later freed) on the caller's
                    if t is not None:
                        t.record_stream(main)
    rank = lr.view(P, 2).mean(dim=1)
    out.update(kl=kl, ap=ap, rank=rank, l1=l1)
    # weighted sum of the four per-pair losses, mean over the pairs: one small product instead of eight elementwise kernels
    wvec = _weight_vector((w['ap'], w['depth'], w['intra'], w['kl']), inv_p, dev)
    out['total'] = (torch.stack((ap, l1, rank, kl)) * wvec).sum()

    if backward:
        # scatter the keypoint gradients back into the token maps (K3 backward)
        gk = gkf.reshape(P, 2, K, Ch)
        cont = (K * Cd, Cd, 1)
        gg1, gg2 = gmaps['g1'], gmaps['g2']
        # descriptor gradients (through the normalisation); when the depth features come from the same map their
        # gradient shares the scatter
        ops.sample_bwd_raw(gd1, cont, d1, ostr, inv1, kp1, (lays['g1'][0], P, K, Cd), geom, True, gg1, lays['g1'][5],
                           gk[:, 0] if shared1 else None, pstr if shared1 else (0, 0, 0))
        ops.sample_bwd_raw(gd2, cont, d2, ostr, inv2, kp2, (lays['g2'][0], P, K, Cd), geom, True, gg2, lays['g2'][5],
                           gk[:, 1] if shared2 else None, pstr if shared2 else (0, 0, 0))
        grads_h = {}
        for name, shared, view, kp in (('h1', shared1, 0, kp1), ('h2', shared2, 1, kp2)):
            if not shared:
                gh = gmaps[name]
                ops.sample_bwd_raw(gk[:, view], pstr, None, (0, 0, 0), None, kp, (lays[name][0], P, K, Ch), geom, False,
                                   gh, lays[name][5])
                grads_h[name] = gh
        out['grads'] = dict(f1=gf1, f2=gf2, g1=gg1, g2=gg2, head=gparams, **grads_h)
    return out


_CONSTS = {}


def _const_vector(n, value, dev):
    """Cached constant fp32 vector (the per-set loss weights): no fill kernel per step, and none inside a captured graph."""
    key = ('full', int(n), float(value), str(dev))
    t = _CONSTS.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            return torch.full((n,), value, dtype=_F32, device=dev)      # never cache graph-pool memory
        t = _CONSTS[key] = torch.full((n,), value, dtype=_F32, device=dev)
    return t


def _weight_vector(weights, scale, dev):
    key = ('w', tuple(float(x) for x in weights), float(scale), str(dev))
    t = _CONSTS.get(key)
    if t is None:
        vals = [float(x) * float(scale) for x in weights]
        if torch.cuda.is_current_stream_capturing():      # no host -> device copy inside a capture: device-side fills
            return torch.cat([torch.full((1, 1), v, dtype=_F32, device=dev) for v in vals])
        t = _CONSTS[key] = torch.tensor(vals, dtype=_F32, device=dev)[:, None]
    return t


class GraphedStep:
    """CUDA-graph replay of ``distillation_step`` for a batch that lives in fixed device buffers.

    The step is ~35 kernel launches and a dozen small torch fills / memsets; captured once, a replay issues them as
    one graph launch (no per-launch gaps, no Python between kernels), with the KL, Smooth-AP and depth-ranking pipelines
    as parallel branches (``parallel_branches=True`` unless the caller says otherwise).  ``step = GraphedStep(batch, variant=...,
    grid=...)``, then refresh the contents of ``batch``'s tensors in place and call ``step()``: it returns the same
    dict of (static) output tensors every time.  Capture allocates the outputs and workspaces from the graph's
    private pool, so the addresses baked into the TMA descriptors stay valid for every replay.
    """

    def __init__(self, batch, **step_kwargs):
        self.batch = batch
        step_kwargs.setdefault('parallel_branches', True)      # parallel graph branches: 3.32 -> 3.24 ms at cfg2
        self.kwargs = step_kwargs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):      # lazy one-time settings (kernel attributes, allocator pools) happen outside the capture
                distillation_step(batch, **step_kwargs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = distillation_step(batch, **step_kwargs)

    def __call__(self):
        self.graph.replay()
        return self.out


class PinnedBatch:
    """All tensors of one host batch in ONE pinned arena, so that the upload is a single copy.

    ``PinnedBatch(batch)`` lays the (CPU) tensors of ``batch`` out back to back (256-byte aligned) in one pinned
    uint8 buffer; non-tensor entries (e.g. the head parameters that live on the device) are passed through.
    ``views(arena)`` re-creates the dict as typed views into any uint8 buffer of the same layout -- the host arena
    or its device copy.  A data loader fills ``host_views()`` in place for the next step instead of allocating."""

    ALIGN = 256

    def __init__(self, batch):
        self.meta = []          # (key, dtype, shape, offset, nbytes)
        self.extra = {}
        off = 0
        for k, v in batch.items():
            if torch.is_tensor(v) and v.device.type == 'cpu':
                n = v.numel() * v.element_size()
                self.meta.append((k, v.dtype, tuple(v.shape), off, n))
                off = (off + n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            else:
                self.extra[k] = v
        self.nbytes = off
        self.arena = torch.empty(max(off, 1), dtype=torch.uint8).pin_memory()
        hv = self.views(self.arena)
        for k, v in batch.items():
            if k not in self.extra:
                hv[k].copy_(v)

    def views(self, arena):
        out = dict(self.extra)
        for k, dtype, shape, off, n in self.meta:
            out[k] = arena[off:off + n].view(dtype).view(shape)
        return out

    def host_views(self):
        return self.views(self.arena)


class DevicePrefetcher:
    """Double-buffered host -> device staging for batches that live in pinned host memory.

    ``for dev_batch in DevicePrefetcher(batches, device): ...`` uploads batch i+1 on a side stream while
    the caller's stream works on batch i, so a step costs max(copy, compute) instead of their sum.  Every
    tensor of every batch is still copied exactly once (non-tensor entries such as the head parameters
    are passed through).  A ``PinnedBatch`` goes over as ONE copy into a device arena that is reused every ``depth``
    steps (no allocation in the steady state); a plain dict of pinned tensors is copied tensor by tensor.  The
    consumer stream waits on the copy event before it touches a batch, and the copy stream waits until the consumer
    has finished with the buffer it is about to overwrite.
    """

    def __init__(self, batches, device, depth=2):
        self.batches = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.queue = []          # [(device batch, ready event, slot)]
        self.slot_done = [None] * depth    # consumer-side event of the last batch that used a slot (guards its reuse)
        self.arenas = []         # device arenas for PinnedBatch uploads, one per slot
        self.n_enqueued = 0

    def _enqueue(self):
        try:
            host = next(self.batches)
        except StopIteration:
            return False
        slot = self.n_enqueued % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self.slot_done[slot] is not None:      # the consumer must be done with the batch that lived in this slot
                self.copy_stream.wait_event(self.slot_done[slot])
            if isinstance(host, PinnedBatch):
                if len(self.arenas) <= slot:
                    self.arenas.append(torch.empty(max(host.nbytes, 1), dtype=torch.uint8, device=self.device))
                elif self.arenas[slot].numel() < host.nbytes:
                    self.arenas[slot] = torch.empty(host.nbytes, dtype=torch.uint8, device=self.device)
                arena = self.arenas[slot]
                arena[:host.arena.numel()].copy_(host.arena, non_blocking=True)
                dev = host.views(arena)
            else:
                dev = {}
                for k, v in host.items():
                    dev[k] = v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.n_enqueued += 1
        self.queue.append((dev, ev, slot))
        return True

    def __iter__(self):
        for _ in range(self.depth):
            self._enqueue()
        while self.queue:
            dev, ev, slot = self.queue.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for v in dev.values():
                if torch.is_tensor(v):
                    v.record_stream(cur)
            yield dev
            done = torch.cuda.Event()
            done.record(cur)
            self.slot_done[slot] = done
            self._enqueue()
