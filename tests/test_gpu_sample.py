"""GPU parity of the bilinear token sampler (forward + backward) against goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import bodies, functions as ofn, synth
from helpers import assert_grad_close

pytestmark = pytest.mark.gpu
T = torch.as_tensor


def test_interpolate_features_golden(golden):
    from gd3.compat import functions as fn
    g = golden('helpers.npz')
    fmap, pts = T(g['interp/fmap']).cuda(), T(g['interp/pts']).cuda()
    for nrm in (False, True):
        out = fn.interpolate_features(fmap, pts, h=9 * 14, w=13 * 14, normalize=nrm)
        np.testing.assert_allclose(out.cpu().numpy(), g[f'interp/out_norm{int(nrm)}'], rtol=0, atol=2e-6)
    out = fn.interpolate_features(T(g['interp16/fmap']).cuda(), T(g['interp16/pts']).cuda(), h=320, w=480,
                                  normalize=False, patch_size=16, stride=16)
    np.testing.assert_allclose(out.cpu().numpy(), g['interp16/out'], rtol=0, atol=2e-6)


@pytest.mark.parametrize('normalize', [False, True])
def test_interpolate_features_backward(normalize):
    from gd3.compat import functions as fn
    gen = synth._gen(9)
    fmap = torch.randn(2, 40, 7, 11, generator=gen)
    pts = torch.stack([torch.rand(2, 33, generator=gen) * 11 * 14, torch.rand(2, 33, generator=gen) * 7 * 14], -1)
    w = torch.randn(2, 40, 33, generator=gen)
    a = fmap.clone().requires_grad_(True)
    (ofn.interpolate_features(a, pts, 7 * 14, 11 * 14, normalize=normalize) * w).sum().backward()
    b = fmap.cuda().requires_grad_(True)
    out = fn.interpolate_features(b, pts.cuda(), 7 * 14, 11 * 14, normalize=normalize)
    (out * w.cuda()).sum().backward()
    assert_grad_close(b.grad, a.grad, cos_min=0.99999, name='fmap', norm_rtol=1e-3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('layers', [0, 4])
def test_sample_tokens_vs_oracle(dtype, layers):
    from gd3 import ops
    ph, pw, C, K, P = 16, 20, 96, 57, 3
    gen = synth._gen(31 + layers)
    shape = (P, ph * pw, C) if layers == 0 else (layers, P, ph * pw, C)
    tok = synth.bf16_round(torch.randn(*shape, generator=gen))
    kp = torch.stack([torch.randint(0, pw * 14, (P, K), generator=gen),
                      torch.randint(0, ph * 14, (P, K), generator=gen)], -1).float()
    w = torch.randn(P, K, C, generator=gen)
    for normalize in (False, True):
        a = tok.clone().requires_grad_(True)
        if layers == 0:
            ref = bodies.sample_tokens(a, ph, pw, kp, normalize=False)
        else:
            ref = torch.stack([bodies.sample_tokens(a[l], ph, pw, kp, normalize=False) for l in range(layers)]).mean(0)
        if normalize:
            ref = torch.nn.functional.normalize(ref, p=2, dim=-1)
        (ref * w).sum().backward()
        b = tok.to('cuda', dtype).requires_grad_(True)
        out = ops.sample_tokens(b, (ph, pw), kp.cuda(), normalize=normalize)
        (out * w.cuda()).sum().backward()
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=0, atol=3e-6)
        assert_grad_close(b.grad, a.grad, cos_min=0.9999, name='tokens', norm_rtol=1e-2)


def test_sample_tokens_empty():
    from gd3 import ops
    tok = torch.randn(2, 16, 8, device='cuda', requires_grad=True)
    out = ops.sample_tokens(tok, (4, 4), torch.zeros(2, 0, 2, device='cuda'))
    assert out.shape == (2, 0, 8)
    out.sum().backward()
    assert float(tok.grad.abs().max()) == 0.0


def test_small_helpers_golden(golden):
    from gd3.compat import functions as fn
    g = golden('helpers.npz')
    out = fn.extract_kp_depth(T(g['kpdepth/depth']).cuda(), T(g['kpdepth/kp']).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), g['kpdepth/out'], rtol=0, atol=2e-6)
    kp = T(g['kpmask/kp']).cuda()
    assert (fn.get_patch_mask_from_kp_tensor(kp, 168, 224, 14).cpu().numpy() == g['kpmask/out']).all()
    assert (fn.get_patch_mask_from_kp_tensor(kp - 1000, 168, 224, 14).cpu().numpy() == g['kpmask/out_empty']).all()
    xs = T(g['sigmoid/x']).cuda()
    np.testing.assert_allclose(fn.sigmoid(xs, 0.01).cpu().numpy(), g['sigmoid/y_t001'], rtol=1e-5, atol=0)
    cost, m1, m2 = T(g['mpc/cost']).cuda(), T(g['mpc/m1']).cuda(), T(g['mpc/m2']).cuda()
    np.testing.assert_allclose(fn.get_masked_patch_cost(cost, m1).cpu().numpy(), g['mpc/rownorm'], rtol=1e-6)
    np.testing.assert_allclose(fn.get_masked_patch_cost(cost, m1, use_softmax=True, temperature=0.5).cpu().numpy(),
                               g['mpc/softmax'], rtol=1e-6)
    np.testing.assert_allclose(fn.get_masked_patch_cost(cost, m1, m2).cpu().numpy(), g['mpc/rownorm_m2'], rtol=1e-6)
    _, idx = fn.filter_kp_by_conf(T(g['conf/kp']).cuda(), T(g['conf/mask']).cuda())
    assert (idx.cpu().numpy() == g['conf/idx']).all()


def test_kp_prepare_batched_matches_oracle():
    """gd3_kp_prepare over several pairs at once against the per-pair oracle helpers, incl. border and outside keypoints."""
    from gd3 import _lib
    from oracle import functions as ofn
    g = torch.Generator().manual_seed(12)
    P, K, H, W, p = 5, 200, 168, 224, 14
    kp = torch.stack([torch.randint(-20, W + 20, (P, K), generator=g), torch.randint(-20, H + 20, (P, K), generator=g)], -1).float()
    kp[0, :4] = torch.tensor([[0., 0.], [W - 1., H - 1.], [W, 5.], [5., H]])
    depth = torch.rand(P, H, W, generator=g) * 5
    mask, kd = _lib.kp_prepare(kp.cuda(), H, W, patch_size=p, depth=depth.cuda(), window=3)
    inside = kp.clone()
    inside[..., 0].clamp_(0, W - 1)
    inside[..., 1].clamp_(0, H - 1)
    _, kd_in = _lib.kp_prepare(inside.cuda(), H, W, depth=depth.cuda(), window=3)
    for i in range(P):
        assert (mask[i].cpu() == ofn.get_patch_mask_from_kp_tensor(kp[i], H, W, p)).all(), i
        want = ofn.extract_kp_depth(depth[i], inside[i:i + 1])
        assert torch.allclose(kd_in[i:i + 1].cpu(), want, rtol=1e-6, atol=1e-6), i
    # one depth map shared by all pairs (stride 0) and a 5 x 5 window
    _, kd5 = _lib.kp_prepare(inside.cuda(), H, W, depth=depth[0].cuda(), window=5)
    for i in range(P):      # the reference (and the oracle) gather one pair at a time
        want5 = ofn.extract_kp_depth(depth[0], inside[i:i + 1], window_size=5)
        assert torch.allclose(kd5[i:i + 1].cpu(), want5, rtol=1e-6, atol=1e-6), i


def test_kp_depth_fractional_and_outside_keypoints():
    """extract_kp_depth gathers at (y * W + x).long() computed in fp32 (utils/functions.py:366-369): fractional
    keypoints follow the flat index, not per-axis truncation; an index outside the map (the reference's gather raises)
    comes back as NaN."""
    from gd3 import _lib
    from oracle import functions as ofn
    g = torch.Generator().manual_seed(13)
    P, K, H, W = 3, 300, 96, 131
    kp = torch.stack([torch.rand(P, K, generator=g) * (W - 1), torch.rand(P, K, generator=g) * (H - 2)], -1)
    depth = torch.rand(P, H, W, generator=g) * 5
    _, kd = _lib.kp_prepare(kp.cuda(), H, W, depth=depth.cuda(), window=3)
    for i in range(P):
        want = ofn.extract_kp_depth(depth[i], kp[i:i + 1])
        assert torch.equal(kd[i:i + 1].cpu(), want) or torch.allclose(kd[i:i + 1].cpu(), want, rtol=1e-6, atol=1e-6), i
    out = torch.tensor([[[0., float(H)], [-3., -1.], [5., 5.]]])
    _, kd_out = _lib.kp_prepare(out.cuda(), H, W, depth=depth[:1].cuda(), window=3)
    assert torch.isnan(kd_out[0, :2]).all() and torch.isfinite(kd_out[0, 2])
