"""Extra blocks of the bench line (bench.py `extra`): the other BASELINE.json configurations on the same GPU(s).

  cfg3         mast3r fast_nn reciprocal NN on 8192 x 8192 x 24 descriptors (kernel time, call time, FMA fraction)
  cfg4         VGGT-teacher ScanNet++ shape (ViT-L/14 518 px: 1369 tokens x 1024, K = 300), 64 pairs on ONE GPU
  cfg4_strong  (N > 1) the same 64 pairs sharded over the N ranks with gd3.dist.shard_batch: strong scaling
  cfg5         (N > 1, or --cfg5) a full fine-tune step: random-init CLIP-shaped ViT-L/14 forward + backward in torch,
               the three distillation losses through the gd3 autograd ops, bucketed NCCL all-reduce of the gradients
               overlapped with the backward (gd3.dist.BucketedGradAllReduce), grad-clip 1.0, AdamW
               (src/main.py:147-159, src/finetune_timm_mast3r.py:683-689)
Inputs of cfg4 / cfg5 are generated on the device (they are not part of any timed region); cfg3 uses the
exactly-representable descriptor set of SURVEY 8-d so that `bit_exact` is meaningful.
"""
import os
import sys
import time

import torch

_CFG3 = {}


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPU cores NVML reports as local to GPU ``local`` (so that the pinned host buffers it
    allocates afterwards are first-touched on that NUMA node).  Returns a short description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f'rank bound to the {len(allowed)} cores local to its GPU (NVML cpu affinity, first {allowed[0]})'
        return 'NVML reported no local cores inside the allowed set; affinity unchanged'
    except Exception as exc:          # no NVML / not permitted: keep going, the line records it
        return f'affinity unchanged ({exc!r})'


def _gen(seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return g


# ---------------------------------------------------------------------------------------------
# cfg3: reciprocal NN
# ---------------------------------------------------------------------------------------------
def cfg3_block(dev, clocks_mhz=1965.0):
    from gd3 import _lib
    from gd3.compat import fast_nn
    n, dim = 8192, 24
    g = torch.Generator().manual_seed(301)
    A = torch.randint(-8, 9, (n, dim), generator=g).float() / 8.0      # entries k / 8: every dot product is exact in fp32
    B = torch.randint(-8, 9, (n, dim), generator=g).float() / 8.0
    B[torch.randperm(n, generator=g)[:64]] = A[torch.randperm(n, generator=g)[:64]]     # duplicates -> ties
    _CFG3['A'], _CFG3['B'] = A, B
    Ad, Bd = A.to(dev), B.to(dev)
    for _ in range(3):
        _lib.reciprocal_nn(Ad, Bd, dist='dot')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    e0.record()
    for _ in range(iters):
        _lib.reciprocal_nn(Ad, Bd, dist='dot')
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    An, Bn = A.numpy(), B.numpy()
    fast_nn.bruteforce_reciprocal_nns(An, Bn, device='cuda', dist='dot', block_size=2 ** 13)
    t0 = time.perf_counter()
    for _ in range(10):
        a, b = fast_nn.bruteforce_reciprocal_nns(An, Bn, device='cuda', dist='dot', block_size=2 ** 13)
    call_ms = (time.perf_counter() - t0) / 10 * 1e3
    _CFG3['a'], _CFG3['b'] = a, b
    flops = 2.0 * n * n * dim
    peak32 = 148 * 128 * 2 * clocks_mhz * 1e6 / 1e12
    return dict(shape='8192 x 8192 x 24, dist=dot', gpu_ms_per_call=round(ms, 4),
                tflops_fp32=round(flops / (ms * 1e-3) / 1e12, 2), flops='2 * nA * nB * dim',
                frac_of_fp32_fma_peak=round(flops / (ms * 1e-3) / 1e12 / peak32, 4),
                call_ms_numpy_in_numpy_out=round(call_ms, 3),
                note='gpu_ms: device-resident descriptors, kernels only (CUDA events over 50 calls); call_ms: '
                     'gd3.compat.fast_nn.bruteforce_reciprocal_nns with host arrays in and out (H2D + one D2H)')


def cfg3_cpu_check():
    """CPU leg of cfg3 (part of bench.py's cpu_baseline): the oracle's bruteforce_reciprocal_nns on the same
    descriptors, timed, and compared index by index with what the GPU returned."""
    from oracle import fast_nn as onn
    A, B = _CFG3['A'], _CFG3['B']
    torch.set_num_threads(os.cpu_count() or 1)
    onn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot', block_size=2 ** 13)
    t0 = time.perf_counter()
    for _ in range(3):
        ra, rb = onn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot', block_size=2 ** 13)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    return dict(cpu_oracle_ms=round(cpu_ms, 2), cpu_threads=torch.get_num_threads(),
                bit_exact=bool((_CFG3['a'] == ra).all() and (_CFG3['b'] == rb).all()))


# ---------------------------------------------------------------------------------------------
# cfg4: device-generated batch of the VGGT-teacher shape
# ---------------------------------------------------------------------------------------------
def make_device_batch(P, N, C, K, grid, variant, dev, seed=4000, heads=4):
    """Synthetic batch of SURVEY 8-d's distributions, generated on the device (same keys as bench_common.make_batch,
    teacher volumes in the producers' packed form)."""
    from gd3 import ops
    g = _gen(seed, dev)
    ph, pw = grid
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)
    batch = dict(f1=rn(P, N, C).to(torch.bfloat16), f2=rn(P, N, C).to(torch.bfloat16))
    for d in ('12', '21'):
        vol = torch.empty(P, N, N, device=dev)
        for p in range(P):
            acc = torch.zeros(N, N, device=dev)
            nh = heads if variant == 'vggt' else 1
            for _ in range(nh):
                logits = 4.0 * rn(N, N)
                logits[torch.arange(N, device=dev), torch.randperm(N, generator=g, device=dev)] += 6.0
                acc += torch.softmax(logits, dim=-1)
            vol[p] = acc / nh
            if variant == 'mast3r':
                vol[p, :, 0] = vol[p].min()
        batch['t' + d], batch['ts' + d] = ops.pack_teacher(vol)
        del vol
    batch['m1'] = torch.rand(P, N, generator=g, device=dev) < 0.6
    batch['m2'] = torch.rand(P, N, generator=g, device=dev) < 0.6
    base = rn(1, 1, C)
    g1 = base + 0.12 * rn(P, N, C)
    batch['g1'] = g1.to(torch.bfloat16)
    batch['g2'] = (g1 + 0.03 * rn(P, N, C)).to(torch.bfloat16)
    W, H = pw * 14, ph * 14
    kp1 = torch.stack([torch.randint(3, W - 3, (P, K), generator=g, device=dev),
                       torch.randint(3, H - 3, (P, K), generator=g, device=dev)], -1).float()
    jit = torch.randint(-4, 5, (P, K, 2), generator=g, device=dev).float()
    kp2 = kp1 + jit
    kp2[..., 0].clamp_(3, W - 4)
    kp2[..., 1].clamp_(3, H - 4)
    batch['kp1'], batch['kp2'] = kp1, kp2
    p1 = torch.rand(P, K, 3, generator=g, device=dev)
    batch['p3d1'], batch['p3d2'] = p1, p1 + 0.02 * rn(P, K, 3)
    batch['dep1'] = torch.rand(P, K, generator=g, device=dev) * 4.5 + 0.5
    batch['dep2'] = torch.rand(P, K, generator=g, device=dev) * 4.5 + 0.5
    k1, k2 = C ** -0.5, 128 ** -0.5
    u = lambda *s: torch.rand(*s, generator=g, device=dev) * 2 - 1
    batch['head'] = dict(W1=u(128, C) * k1, b1=u(128) * k1, gamma=1.0 + 0.1 * rn(128), beta=0.1 * rn(128),
                         w2=u(1, 128) * k2, b2=u(1) * k2, use_tanh=True, ln_eps=1e-5)
    return batch


def _graph_time(batch, variant, grid, steps, world, local, sync_world=True):
    from bench import timed_replays
    from gd3 import pipeline
    gs = pipeline.GraphedStep(batch, variant=variant, grid=grid, backward=True)
    for _ in range(3):
        gs()
    ms, _, _ = timed_replays(gs, steps, world if sync_world else 1, local, sample_clocks=False)
    return ms, gs


def cfg4_blocks(args, rank, world, local, dev, peaks):
    from bench import WORKLOADS, gemm_flops_per_step
    from gd3 import _lib, dist as gdist, pipeline
    c = WORKLOADS['cfg4']
    P, N, C, K, grid, variant = c['P'], c['N'], c['C'], c['K'], c['grid'], c['variant']
    steps = max(5, min(args.steps, 20))
    full = make_device_batch(P, N, C, K, grid, variant, dev)
    out = {}
    # ---- all 64 pairs on one GPU (every rank runs the same problem on its own GPU; max over ranks) ----
    ms1, gs = _graph_time(full, variant, grid, steps, world, local)
    blk = dict(workload=f"cfg4: {c['desc']}, {P} pairs on ONE GPU", ms_per_step=round(ms1, 4),
               value=round(P / (ms1 * 1e-3), 2), unit='pairs/s', steps=steps, launch_mode='CUDA-graph replay',
               data='synthetic, generated on the device; teacher volumes packed (fp16 x 1024 + row statistics)')
    del gs
    # per-kernel shares from a short eager, profiled run
    _lib.profile_enable(True)
    _lib.profile_read()
    for _ in range(3):
        pipeline.distillation_step(full, variant=variant, grid=grid, backward=True)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    tot = sum(ms for _, ms in prof.values()) or 1.0
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]
    blk['kernels'] = {k: dict(us_per_step=round(ms / 3 * 1e3, 1), share=round(ms / tot, 4)) for k, (cnt, ms) in top}
    fl = gemm_flops_per_step(c, P)
    if 'kl_grad_gemm' in prof:
        cnt, ms = prof['kl_grad_gemm']
        ach = fl['kl_grad_gemm'] * 3 / cnt / (ms / cnt * 1e-3) / 1e12
        blk['kl_grad_gemm'] = dict(us_per_launch=round(ms / cnt * 1e3, 1), tflops=round(ach, 1),
                                   frac_of_burst_bf16_peak=round(ach / peaks['bf16_tflops'], 4))
    out['cfg4'] = blk
    # ---- strong scaling: the SAME 64 pairs sharded over the ranks, no collective in the loss path ----
    if world > 1:
        shard = gdist.shard_batch(full, rank, world)       # views: the packed teacher rows keep their 16-byte padding
        del full
        msn, gs = _graph_time(shard, variant, grid, steps, world, local)
        out['cfg4_strong'] = dict(
            workload=f'cfg4, {P} pairs in total, {shard["f1"].shape[0]} per GPU (gd3.dist.shard_batch)',
            n_gpus=world, ms_per_step=round(msn, 4), value=round(P / (msn * 1e-3), 2), unit='pairs/s',
            single_gpu_ms_per_step=round(ms1, 4), speedup=round(ms1 / msn, 3), efficiency=round(ms1 / msn / world, 4),
            scaling='strong', note='time = max over ranks of the CUDA-event time of K graph replays; the single-GPU '
                                   'figure is the same batch unsharded, measured on every GPU of this run (max)')
        del gs
    return out


# ---------------------------------------------------------------------------------------------
# cfg5: full fine-tune step with the gradient all-reduce
# ---------------------------------------------------------------------------------------------
class DepthHead(torch.nn.Module):
    """Shape of DepthAwareFeatureFusion.fusion_layer (+ tanh), utils/model.py:100-105,122-127."""

    def __init__(self, dim, hidden=128):
        super().__init__()
        self.fusion_layer = torch.nn.Sequential(torch.nn.Linear(dim, hidden), torch.nn.LayerNorm(hidden),
                                                torch.nn.GELU(), torch.nn.Linear(hidden, 1))
        self.use_tanh = True


def allreduce_bench(nbytes, world, dev, iters=10):
    import torch.distributed as dist
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=dev)
    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return dict(payload_mb=round(nbytes / 1e6, 1), ms=round(ms, 4),
                bus_gbs=round(2.0 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9, 1))


def cfg5_block(args, rank, world, local, dev, pairs=4, K=128, steps=11, warmup=4, bucket_mb=256):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from bench import barrier, max_over_ranks
    from gd3 import dist as gdist, ops
    torch.manual_seed(0)                       # identical initial weights on every rank (what DDP's broadcast gives)
    vcfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                            image_size=224, patch_size=14)
    vit = CLIPVisionModel(vcfg).to(dev)
    head = DepthHead(1024).to(dev)
    params = [p for p in list(vit.parameters()) + list(head.parameters())]
    n_params = sum(p.numel() for p in params)
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4, fused=True)      # src/finetune_timm_mast3r.py:683-689
    P, N, grid = pairs, 256, (16, 16)
    g = _gen(5000 + rank, dev)
    imgs = torch.randn(2 * P, 3, 224, 224, generator=g, device=dev)
    aux = make_device_batch(P, N, 8, K, grid, 'mast3r', dev, seed=5100 + rank)     # teacher volumes, masks, keypoints, depths
    t12, t21 = (aux['t12'], aux['ts12']), (aux['t21'], aux['ts21'])
    depths = torch.stack([aux['dep1'], aux['dep2']], 1).reshape(2 * P, K)
    w_rank = torch.full((2 * P,), 0.5 / P, device=dev)
    w_l1 = torch.full((P,), 1.0 / P, device=dev)

    def train_step(red):
        red.reset()
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            o = vit(pixel_values=imgs, output_hidden_states=True)
        tok = o.last_hidden_state[:, 1:, :]                                  # (2P, 256, 1024) patch tokens
        mid = torch.stack(o.hidden_states[12:16])[:, :, 1:, :]               # 4 mid blocks for the depth features
        f1, f2 = tok[0::2], tok[1::2]
        kl = ops.cost_volume_kl(f1, f2, t12, t21, aux['m1'], aux['m2'], variant='mast3r').mean()
        d1 = ops.sample_tokens(f1, grid, aux['kp1'], normalize=True)
        d2 = ops.sample_tokens(f2, grid, aux['kp2'], normalize=True)
        ap = ops.smooth_ap(d1, d2, aux['p3d1'], aux['p3d2'], variant='mast3r').mean()
        kf1 = ops.sample_tokens(mid[:, 0::2], grid, aux['kp1'])
        kf2 = ops.sample_tokens(mid[:, 1::2], grid, aux['kp2'])
        feats = torch.stack([kf1, kf2], 1).reshape(2 * P, K, kf1.shape[-1])
        dl, _, _ = ops.depth_head_loss(head, feats, depths, w_rank=w_rank, w_l1=w_l1)
        loss = kl + ap + dl
        loss.backward()
        norm = red.finish(clip_norm=1.0)                                     # gradient_clip_val=1.0, src/main.py:158
        opt.step()
        return loss.detach(), norm

    def timed(red, n):
        # the step is ~2000 small torch launches per rank on a shared host: single steps jitter by +-10 ms with 8 ranks,
        # so every step is timed on its own and the MEDIAN step is reported (max over ranks of the medians)
        for _ in range(warmup):
            train_step(red)
        barrier(world)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            loss, norm = train_step(red)
            evs[i + 1].record()
        barrier(world)
        per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n))
        return max_over_ranks(per_step[n // 2], world), float(loss), float(norm)

    red = gdist.BucketedGradAllReduce(params, bucket_bytes=bucket_mb * 1024 * 1024)
    ms_full, loss, norm = timed(red, steps)
    red.remove()
    # the same step with the collective switched off (every rank keeps its local gradients): exposed communication
    red_local = gdist.BucketedGradAllReduce(params, bucket_bytes=bucket_mb * 1024 * 1024)
    red_local.world = 1
    ms_local, _, _ = timed(red_local, steps)
    red_local.remove()
    blk = dict(workload=f'cfg5: full fine-tune step, CLIP-shaped ViT-L/14 224 px (random init, {n_params / 1e6:.0f} M '
                        f'parameters, bf16 autocast), {P} pairs per GPU, KL + Smooth-AP + depth losses through the '
                        f'gd3 autograd ops, bucketed NCCL all-reduce overlapped with backward, clip 1.0, fused AdamW',
               n_gpus=world, pairs_per_gpu=P, step_ms=round(ms_full, 3), step_ms_without_allreduce=round(ms_local, 3),
               exposed_allreduce_ms=round(ms_full - ms_local, 3), value=round(world * P / (ms_full * 1e-3), 2),
               unit='pairs/s', gradient_payload_mb=round(red.payload_bytes / 1e6, 1), buckets=len(red.buckets),
               bucket_mb=bucket_mb,
               loss=round(loss, 5), grad_norm=round(norm, 5), steps=steps,
               timing='median of the per-step CUDA-event times (max over ranks); the step is host-launch-bound and single steps jitter')
    if world > 1:
        blk['allreduce_full_finetune'] = allreduce_bench(red.payload_bytes, world, dev)
        blk['allreduce_lora_set'] = allreduce_bench(int(25.6e6) // 4 * 4, world, dev)     # the reference's trainable set
    del vit, opt, red, red_local
    torch.cuda.empty_cache()
    return blk


def run(args, rank, world, local, dev, peaks):
    out = {}
    if rank == 0:
        out['cfg3'] = cfg3_block(dev)
    out.update(cfg4_blocks(args, rank, world, local, dev, peaks))
    if world > 1 or getattr(args, 'cfg5', False):
        try:
            out['cfg5'] = cfg5_block(args, rank, world, local, dev)
        except Exception as exc:
            print(f'[bench] cfg5 failed on rank {rank}: {exc!r}', file=sys.stderr)
            out['cfg5'] = dict(error=repr(exc))
    return out
