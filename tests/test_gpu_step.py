"""GPU parity of one whole fused distillation step (KL + Smooth-AP + ranking + L1, fwd + bwd, batched)
against the CPU oracle's per-pair reference flow."""
import pytest
import torch

import bench_common
from helpers import assert_grad_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('variant,cfg', [
    ('mast3r', dict(N=256, C=384, K=128, grid=(16, 16), P=3)),
    ('vggt', dict(N=15 * 17, C=200, K=77, grid=(15, 17), P=2)),
])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_distillation_step_vs_oracle(variant, cfg, dtype):
    from gd3 import pipeline
    cfg = dict(cfg, variant=variant)
    batch = bench_common.make_batch(cfg, cfg_id=1)
    want = bench_common.oracle_step(batch, cfg)
    dev = bench_common.to_device(batch, 'cuda', feature_dtype=dtype)
    out = pipeline.distillation_step(dev, variant=variant, grid=cfg['grid'], backward=True, pairs_per_group=2)
    torch.cuda.synchronize()
    for k in ('kl', 'ap', 'rank', 'l1'):
        got, ref = out[k].float().cpu(), want[k]
        err = ((got - ref).abs() / ref.abs().clamp_min(1e-12)).max().item()
        assert err <= 1e-3, (k, got, ref)
    assert abs(out['total'].item() - want['total'].item()) <= 1e-3 * abs(want['total'].item())
    for k in ('f1', 'f2', 'g1', 'g2', 'head'):
        assert_grad_close(out['grads'][k], want['grads'][k], name=k, norm_rtol=3e-2)
    fwd = pipeline.distillation_step(dev, variant=variant, grid=cfg['grid'], backward=False)
    assert 'grads' not in fwd
    assert abs(fwd['total'].item() - out['total'].item()) <= 1e-5 * abs(out['total'].item())


def test_step_with_separate_depth_feature_maps():
    """Descriptors from one map, depth-head features from a 4-block stack of other maps (their mean), as the reference's
    training_step does (refine_conv output vs get_intermediate_feature, src/finetune_timm_mast3r.py:271-277,307-313)."""
    from gd3 import pipeline
    cfg = dict(N=12 * 16, C=96, K=65, grid=(12, 16), P=2, variant='vggt')
    batch = bench_common.make_batch(cfg, cfg_id=1, pair0=3)
    g = torch.Generator().manual_seed(21)
    batch['h1'] = 0.5 * torch.randn(4, cfg['P'], cfg['N'], 72, generator=g)
    batch['h2'] = batch['h1'] + 0.1 * torch.randn(4, cfg['P'], cfg['N'], 72, generator=g)
    from oracle import synth
    batch['head'] = dict(synth.head_params(99, 72), use_tanh=True, ln_eps=1e-5)
    want = bench_common.oracle_step(batch, cfg)
    out = pipeline.distillation_step(bench_common.to_device(batch, 'cuda'), variant='vggt', grid=cfg['grid'])
    torch.cuda.synchronize()
    for k in ('kl', 'ap', 'rank', 'l1'):
        got, ref = out[k].float().cpu(), want[k]
        assert ((got - ref).abs() / ref.abs().clamp_min(1e-12)).max().item() <= 1e-3, (k, got, ref)
    for k in ('g1', 'g2', 'h1', 'h2', 'head'):
        assert out['grads'][k].shape == want['grads'][k].shape, k
        assert_grad_close(out['grads'][k], want['grads'][k], name=k, norm_rtol=3e-2)
    bad = dict(bench_common.to_device(batch, 'cuda'))
    bad['h1'] = bad['h1'][:, :, :-1]
    with pytest.raises(ValueError):
        pipeline.distillation_step(bad, variant='vggt', grid=cfg['grid'])


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_graphed_step_matches_eager():
    """CUDA-graph replay of the fused step: same losses and gradients as the eager call, also after the inputs change."""
    import bench_common
    from gd3 import pipeline
    cfg = dict(N=256, C=384, K=128, P=3, grid=(16, 16), variant='mast3r')
    batch = bench_common.to_device(bench_common.make_batch(cfg, cfg_id=1), 'cuda', feature_dtype=torch.bfloat16)
    eager = pipeline.distillation_step(batch, variant='mast3r', grid=cfg['grid'])
    step = pipeline.GraphedStep(batch, variant='mast3r', grid=cfg['grid'])
    out = step()
    torch.cuda.synchronize()
    for k in ('kl', 'ap', 'rank', 'l1'):
        assert torch.allclose(out[k], eager[k], rtol=1e-5, atol=1e-7), k
    for k in ('f1', 'g1', 'head'):
        # fp32 atomics (token-map scatter, split-K d W1) make the two runs differ in the last bits: compare direction and size
        a, b = out['grads'][k].float().flatten().double(), eager['grads'][k].float().flatten().double()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        assert cos > 1 - 1e-6 and abs(float(a.norm() / b.norm()) - 1) < 1e-4, (k, cos)
    # new contents in the same buffers -> new results from the same graph
    other = bench_common.to_device(bench_common.make_batch(cfg, cfg_id=1, pair0=7), 'cuda', feature_dtype=torch.bfloat16)
    for k, v in other.items():
        if torch.is_tensor(v):
            batch[k].copy_(v)
    want = pipeline.distillation_step(other, variant='mast3r', grid=cfg['grid'])
    got = step()
    torch.cuda.synchronize()
    assert torch.allclose(got['kl'], want['kl'], rtol=1e-5) and torch.allclose(got['rank'], want['rank'], rtol=1e-5)
    assert not torch.allclose(got['kl'], eager['kl'], rtol=1e-3)


def test_parallel_branches_match_serial():
    """The KL / Smooth-AP pipelines on side streams (parallel_branches=True) give the same step as everything on one stream."""
    from gd3 import pipeline
    cfg = dict(N=256, C=384, K=128, P=4, grid=(16, 16), variant='vggt')
    batch = bench_common.to_device(bench_common.make_batch(cfg, cfg_id=1, pair0=11), 'cuda', feature_dtype=torch.bfloat16)
    for _ in range(3):
        a = pipeline.distillation_step(batch, variant='vggt', grid=cfg['grid'], parallel_branches=True)
        b = pipeline.distillation_step(batch, variant='vggt', grid=cfg['grid'], parallel_branches=False)
        torch.cuda.synchronize()
        for k in ('kl', 'ap', 'rank', 'l1'):
            assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-8), k
        for k in ('f1', 'f2', 'g1', 'head'):
            x, y = a['grads'][k].float().flatten().double(), b['grads'][k].float().flatten().double()
            assert float(torch.dot(x, y) / (x.norm() * y.norm())) > 1 - 1e-6, k


def test_step_derives_masks_and_keypoint_depths():
    """Without m1 / m2 / dep1 / dep2 the step builds them from the keypoints and depth maps on the device; the result
    equals the step fed with the oracle helpers' masks (utils/functions.py:375-399) and depths (:348-372)."""
    from oracle import functions as ofn
    from gd3 import pipeline
    cfg = dict(N=256, C=384, K=128, P=3, grid=(16, 16), variant='mast3r')
    H = W = 16 * 14
    batch = bench_common.make_batch(cfg, cfg_id=1)
    g = torch.Generator().manual_seed(5)
    depth_maps = [torch.rand(cfg['P'], H, W, generator=g) * 4 + 0.5 for _ in range(2)]
    explicit = dict(batch)
    for v, dm in (('1', depth_maps[0]), ('2', depth_maps[1])):
        kp = batch['kp' + v]
        explicit['m' + v] = torch.stack([ofn.get_patch_mask_from_kp_tensor(kp[p], H, W, 14) for p in range(cfg['P'])])
        explicit['dep' + v] = torch.cat([ofn.extract_kp_depth(dm[p], kp[p:p + 1]) for p in range(cfg['P'])])
    derived = {k: v for k, v in batch.items() if k not in ('m1', 'm2', 'dep1', 'dep2')}
    derived['depth_map1'], derived['depth_map2'] = depth_maps
    a = pipeline.distillation_step(bench_common.to_device(explicit, 'cuda', feature_dtype=torch.bfloat16),
                                   variant='mast3r', grid=cfg['grid'])
    b = pipeline.distillation_step(bench_common.to_device(derived, 'cuda', feature_dtype=torch.bfloat16),
                                   variant='mast3r', grid=cfg['grid'])
    torch.cuda.synchronize()
    for k in ('kl', 'ap', 'rank', 'l1'):
        assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-7), (k, a[k], b[k])
    assert not torch.allclose(a['kl'], pipeline.distillation_step(
        bench_common.to_device(batch, 'cuda', feature_dtype=torch.bfloat16), variant='mast3r', grid=cfg['grid'])['kl'], rtol=1e-4)


@pytest.mark.parametrize('seed', range(8))
def test_step_random_shapes(seed):
    """Random ragged shapes (token grid, channels not a multiple of 8, K = 1 ...) of the whole step against the
    oracle: same bars as the fixed configurations (tools/probe_shapes.py runs longer sweeps)."""
    import random
    from gd3 import pipeline
    rnd = random.Random(1000 + seed)
    ph, pw = rnd.randint(4, 24), rnd.randint(4, 24)
    cfg = dict(N=ph * pw, C=rnd.choice([64, 72, 96, 100, 128, 200, 384, 388, 520]),
               K=rnd.choice([1, 2, 7, 33, 64, 100, 129, 257]), grid=(ph, pw), P=rnd.randint(1, 4),
               variant=rnd.choice(['mast3r', 'vggt']))
    dtype = rnd.choice([torch.float32, torch.bfloat16])
    batch = bench_common.make_batch(cfg, cfg_id=1, pair0=seed)
    want = bench_common.oracle_step(batch, cfg)
    out = pipeline.distillation_step(bench_common.to_device(batch, 'cuda', feature_dtype=dtype), variant=cfg['variant'],
                                     grid=cfg['grid'], pairs_per_group=rnd.choice([0, 1, 2]))
    torch.cuda.synchronize()
    for k in ('kl', 'ap', 'rank', 'l1'):
        got, ref = out[k].float().cpu(), want[k]
        err = ((got - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
        assert err <= 1e-3, (cfg, dtype, k, got, ref)
    for k in ('f1', 'f2', 'g1', 'g2', 'head'):
        g, r = out['grads'][k].float().cpu(), want['grads'][k]
        if float(r.norm()) < 1e-12 and float(g.norm()) < 1e-9:
            continue
        assert_grad_close(g, r, name=f'{cfg} {k}', norm_rtol=3e-2)


def test_pinned_batch_prefetcher_delivers_every_batch_intact():
    """PinnedBatch (one pinned arena per host batch, one H2D copy) through the double-buffered DevicePrefetcher: every
    step sees exactly its own batch although the two device arenas are reused and the consumer is slow."""
    from gd3 import pipeline
    g = torch.Generator().manual_seed(3)
    steps = 7
    src = [dict(a=torch.randn(1000, 33, generator=g), m=torch.rand(257, generator=g) < 0.5,
                h=torch.randn(64, 8, generator=g).to(torch.float16), k=torch.randint(0, 99, (5, 2), generator=g),
                head=dict(tag=i)) for i in range(steps)]
    packed = [pipeline.PinnedBatch(b) for b in src]
    assert packed[0].arena.is_pinned() and packed[0].nbytes % 256 == 0
    hv = packed[2].host_views()
    assert torch.equal(hv['a'], src[2]['a']) and torch.equal(hv['m'], src[2]['m']) and hv['head']['tag'] == 2
    big = torch.randn(4096, 4096, device='cuda')
    seen = 0
    for i, dev in enumerate(pipeline.DevicePrefetcher(iter(packed), 'cuda')):
        for _ in range(3):
            big = big @ big * 1e-4          # keep the consumer stream busy while the next upload runs
        for k in ('a', 'm', 'h', 'k'):
            assert dev[k].is_cuda and torch.equal(dev[k].cpu(), src[i][k]), (i, k)
        assert dev['head']['tag'] == i
        seen += 1
    assert seen == steps
    # plain dicts of pinned tensors still work (tensor-by-tensor copies)
    plain = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()} for b in src[:3]]
    for i, dev in enumerate(pipeline.DevicePrefetcher(iter(plain), 'cuda')):
        assert torch.equal(dev['a'].cpu(), src[i]['a'])
