"""The CUDA ops against losses and gradients of the reference's own LightningModule loss methods
(``tests/golden/live_bodies.npz``, written by ``oracle/gen_live_bodies.py`` from the live ``calculate_cost_loss`` /
``calculate_matching_loss`` / ``calculate_depth_loss`` of both fine-tune modules).  Bars: BASELINE's (loss relative
error <= 1e-3, gradient cosine >= 0.999)."""
import pytest
import torch

from helpers import assert_grad_close, rel_err
from oracle import losses as olosses
from oracle import synth
from test_oracle_golden import LIVE_TAGS, live_case

pytestmark = pytest.mark.gpu

T = torch.as_tensor


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('tag', LIVE_TAGS)
def test_cost_volume_kl_vs_live_method(golden, tag, dtype):
    from gd3 import ops
    g = golden('live_bodies.npz')
    c = live_case(g, tag)
    # the live method saw these exact values; for the bf16 run they are rounded first, which moves the loss by < 1e-3
    f1 = c['f1'].to(dtype).cuda()[None].requires_grad_(True)
    f2 = c['f2'].to(dtype).cuda()[None].requires_grad_(True)
    kl = ops.cost_volume_kl(f1, f2, c['t12'].cuda()[None], c['t21'].cuda()[None], c['m1'].cuda()[None], c['m2'].cuda()[None],
                            variant=c['variant'])
    kl.sum().backward()
    assert rel_err(kl[0].detach().cpu(), g[f'{tag}/kl']) < (1e-3 if dtype == torch.float32 else 3e-3)
    assert_grad_close(f1.grad[0].float().cpu(), T(g[f'{tag}/grad_f1']), name='f1', norm_rtol=3e-2)
    assert_grad_close(f2.grad[0].float().cpu(), T(g[f'{tag}/grad_f2']), name='f2', norm_rtol=3e-2)


@pytest.mark.parametrize('tag', LIVE_TAGS)
def test_smooth_ap_vs_live_method(golden, tag):
    from gd3 import ops
    g = golden('live_bodies.npz')
    c = live_case(g, tag)
    d1, d2 = c['d1'].cuda().requires_grad_(True), c['d2'].cuda().requires_grad_(True)
    ap = ops.smooth_ap(d1, d2, c['p3d1'].cuda()[None], c['p3d2'].cuda()[None], variant=c['variant'])
    ap.sum().backward()
    assert rel_err(ap[0].detach().cpu(), g[f'{tag}/ap']) < 1e-3
    assert_grad_close(d1.grad.cpu(), T(g[f'{tag}/grad_d1']), name='d1', norm_rtol=3e-2)
    assert_grad_close(d2.grad.cpu(), T(g[f'{tag}/grad_d2']), name='d2', norm_rtol=3e-2)


@pytest.mark.parametrize('tag', LIVE_TAGS)
def test_depth_losses_vs_live_method(golden, tag):
    from gd3 import _lib, ops
    g = golden('live_bodies.npz')
    c = live_case(g, tag)
    head = olosses.DepthHead(c['C'])                       # container of the six parameter tensors only
    synth.load_head(head, c['head_params'])
    head = head.cuda()
    feats = torch.cat([c['kf1'], c['kf2']]).cuda().requires_grad_(True)               # sets (view 1, view 2)
    # keypoint depths from the depth maps on the device (the live method calls extract_kp_depth itself)
    _, kd1 = _lib.kp_prepare(c['kp1'].cuda(), c['dm1'].shape[0], c['dm1'].shape[1], depth=c['dm1'].cuda())
    _, kd2 = _lib.kp_prepare(c['kp2'].cuda(), c['dm2'].shape[0], c['dm2'].shape[1], depth=c['dm2'].cuda())
    one = feats.new_ones(1)
    total, rank, l1 = ops.depth_head_loss(head, feats, torch.cat([kd1, kd2]), depth_threshold=0.05,
                                          w_rank=one.expand(2) * 0.5, w_l1=one)
    total.backward()
    assert rel_err(l1[0].cpu(), g[f'{tag}/l1']) < 1e-3
    assert rel_err((0.5 * (rank[0] + rank[1])).cpu(), g[f'{tag}/rank']) < 1e-3
    assert_grad_close(feats.grad[0].cpu(), T(g[f'{tag}/grad_kf1'])[0], name='kf1', norm_rtol=3e-2)
    assert_grad_close(feats.grad[1].cpu(), T(g[f'{tag}/grad_kf2'])[0], name='kf2', norm_rtol=3e-2)
    fl = head.fusion_layer
    packed = torch.cat([q.grad.reshape(-1) for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight,
                                                     fl[3].bias)]).cpu()
    assert_grad_close(packed, T(g[f'{tag}/grad_head']), name='head', norm_rtol=3e-2)


@pytest.mark.parametrize('tag', ['me0', 'me1'])
def test_me_smooth_ap_vs_live_training_step(golden, tag):
    """'me' variant against the live ``FinetuneTIMM.training_step`` (src/finetune_timm_me.py:191-220)."""
    from gd3 import ops
    g = golden('live_bodies.npz')
    d1, d2 = T(g[f'{tag}/d1']).cuda().requires_grad_(True), T(g[f'{tag}/d2']).cuda().requires_grad_(True)
    ap = ops.smooth_ap(d1, d2, T(g[f'{tag}/p3d1']).cuda()[None], T(g[f'{tag}/p3d2']).cuda()[None], variant='me')
    ap.sum().backward()
    assert rel_err(ap[0].detach().cpu(), g[f'{tag}/ap']) < 1e-3
    assert_grad_close(d1.grad.cpu(), T(g[f'{tag}/grad_d1']), name='d1', norm_rtol=3e-2)
    assert_grad_close(d2.grad.cpu(), T(g[f'{tag}/grad_d2']), name='d2', norm_rtol=3e-2)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('tag', ['sample0', 'sample1'])
def test_sample_tokens_vs_live_feature_getters(golden, tag, dtype):
    """One fused sample of the (4, 1, N, C) layer stack == the live ``get_intermediate_feature`` (4 samples + mean);
    the normalised sample of the final features == the live ``get_feature``.  Forward and token gradients."""
    import numpy as np
    from gd3 import ops
    g = golden('live_bodies.npz')
    gh, gw = (int(v) for v in g[f'{tag}/grid'])
    layers = T(g[f'{tag}/layers']).to(dtype).cuda()[:, None].contiguous().requires_grad_(True)      # (L, P = 1, N, C)
    final = T(g[f'{tag}/final']).to(dtype).cuda()[None].contiguous().requires_grad_(True)
    kp = T(g[f'{tag}/kp']).cuda()
    feat = ops.sample_tokens(layers, (gh, gw), kp, normalize=False)
    desc = ops.sample_tokens(final, (gh, gw), kp, normalize=True)
    tol = dict(rtol=1e-5, atol=2e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(feat.detach().float().cpu().numpy(), g[f'{tag}/feat'], **tol)
    np.testing.assert_allclose(desc.detach().float().cpu().numpy(), g[f'{tag}/desc'], **tol)
    ((feat * T(g[f'{tag}/w_feat']).cuda()).sum() + (desc * T(g[f'{tag}/w_desc']).cuda()).sum()).backward()
    cos_min = 0.99999 if dtype == torch.float32 else 0.999
    assert_grad_close(layers.grad[:, 0].float().cpu(), T(g[f'{tag}/grad_layers']), cos_min=cos_min, name='layers', norm_rtol=3e-2)
    assert_grad_close(final.grad[0].float().cpu(), T(g[f'{tag}/grad_final']), cos_min=cos_min, name='final', norm_rtol=3e-2)


@pytest.mark.parametrize('tag', ['kpmatch0', 'kpmatch1'])
def test_teacher_keypoints_vs_live_method(golden, tag):
    """Device-side reciprocal matching + border / confidence filters == the live
    ``FinetuneMASt3RTIMM.filter_and_match_keypoints`` (src/finetune_timm_mast3r.py:392-469), keypoint for keypoint."""
    from gd3.compat import keypoints
    g = golden('live_bodies.npz')
    feats = dict(desc_1=T(g[f'{tag}/desc1_x8']).float() / 8, desc_2=T(g[f'{tag}/desc2_x8']).float() / 8,
                 conf_1=T(g[f'{tag}/conf1']), conf_2=T(g[f'{tag}/conf2']))
    kp1, kp2, w, h = keypoints.filter_and_match_keypoints(feats, float(g[f'{tag}/min_conf_thr']))
    assert kp1.is_cuda and kp1.dtype == torch.float32
    assert (w, h) == tuple(int(v) for v in g[f'{tag}/wh'])
    assert torch.equal(kp1.cpu(), T(g[f'{tag}/kp1'])) and torch.equal(kp2.cpu(), T(g[f'{tag}/kp2']))
