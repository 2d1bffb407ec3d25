"""GPU parity of the fused Smooth-AP loss (through the C ABI): goldens from the live reference, the CPU
oracle at full size, and the composition sample_tokens -> normalise -> smooth_ap with gradients
flowing back to the token maps."""
import math

import pytest
import torch

from oracle import bodies, synth
from helpers import assert_grad_close, rel_err
from test_oracle_golden import AP_CASES

pytestmark = pytest.mark.gpu
T = torch.as_tensor
VARS = ['mast3r', 'vggt', 'me']


@pytest.mark.parametrize('case', AP_CASES)
def test_smooth_ap_golden(golden, case):
    from gd3 import ops
    g = golden('smooth_ap.npz')
    K, C, seed, var = [int(v) for v in g[f'{case}/meta']]
    d1 = T(g[f'{case}/d1']).cuda()[None].requires_grad_(True)
    d2 = T(g[f'{case}/d2']).cuda()[None].requires_grad_(True)
    p1, p2 = T(g[f'{case}/p1']).cuda()[None], T(g[f'{case}/p2']).cuda()[None]
    loss = ops.smooth_ap(d1, d2, p1, p2, variant=VARS[var])
    loss.sum().backward()
    want = float(g[f'{case}/loss'])
    assert abs(loss[0].item() - want) <= 1e-3 * abs(want) + 1e-7, (loss[0].item(), want)
    assert_grad_close(d1.grad[0], T(g[f'{case}/grad_d1']), name='grad_d1', norm_rtol=3e-2)
    assert_grad_close(d2.grad[0], T(g[f'{case}/grad_d2']), name='grad_d2', norm_rtol=3e-2)


@pytest.mark.parametrize('case', ['small_mast3r', 'odd_vggt', 'mid_me'])
def test_sample_then_smooth_ap_golden(golden, case):
    """Gradients all the way back to the token maps (bilinear scatter + normalise backward)."""
    from gd3 import ops
    g = golden('smooth_ap.npz')
    K, C, seed, var = [int(v) for v in g[f'{case}/meta']]
    g1 = T(g[f'{case}/g1']).cuda()[None].requires_grad_(True)
    g2 = T(g[f'{case}/g2']).cuda()[None].requires_grad_(True)
    kp1, kp2 = T(g[f'{case}/kp1']).cuda()[None], T(g[f'{case}/kp2']).cuda()[None]
    d1 = ops.sample_tokens(g1, (16, 16), kp1, normalize=True)
    d2 = ops.sample_tokens(g2, (16, 16), kp2, normalize=True)
    loss = ops.smooth_ap(d1, d2, T(g[f'{case}/p1']).cuda()[None], T(g[f'{case}/p2']).cuda()[None], variant=VARS[var])
    loss.sum().backward()
    assert rel_err(loss[0].item(), g[f'{case}/loss']) <= 1e-3
    assert_grad_close(g1.grad[0], T(g[f'{case}/grad_g1']), name='grad_g1', norm_rtol=3e-2)
    assert_grad_close(g2.grad[0], T(g[f'{case}/grad_g2']), name='grad_g2', norm_rtol=3e-2)


def make_pair(seed, K, C, N=1024, grid=(32, 32)):
    ph, pw = grid
    g1, g2 = synth.ap_token_maps(seed, N, C)
    kp1 = synth.keypoints(seed + 1, K, pw * 14, ph * 14)
    kp2 = kp1 + (synth.keypoints(seed + 2, K, 9, 9) - 4.0)
    kp2[:, 0].clamp_(3, pw * 14 - 4)
    kp2[:, 1].clamp_(3, ph * 14 - 4)
    p1, p2 = synth.points3d(seed + 3, K)
    d1 = bodies.sample_tokens(g1[None], ph, pw, kp1[None], normalize=True)[0]
    d2 = bodies.sample_tokens(g2[None], ph, pw, kp2[None], normalize=True)[0]
    return d1, d2, p1, p2


@pytest.mark.parametrize('variant,K,C', [('mast3r', 512, 768), ('vggt', 300, 1024)])
def test_smooth_ap_full_size_batched(variant, K, C):
    """BASELINE.json sizes (cfg2: K=512,C=768; cfg4: K=300,C=1024), 3 pairs per call vs the CPU oracle."""
    from gd3 import ops
    pairs = [make_pair(7000 + 10 * p, K, C) for p in range(3)]
    want = []
    for d1, d2, p1, p2 in pairs:
        a = d1.clone().requires_grad_(True)
        b = d2.clone().requires_grad_(True)
        loss = bodies.smooth_ap(a, b, p1, p2, variant)
        ga, gb = torch.autograd.grad(loss, [a, b])
        want.append((float(loss), ga, gb))
    D1, D2, P1, P2 = [torch.stack([q[k] for q in pairs]).cuda() for k in range(4)]
    D1.requires_grad_(True)
    D2.requires_grad_(True)
    loss = ops.smooth_ap(D1, D2, P1, P2, variant=variant)
    loss.sum().backward()
    for p in range(3):
        assert rel_err(loss[p].item(), want[p][0]) <= 1e-3, (p, loss[p].item(), want[p][0])
        assert_grad_close(D1.grad[p], want[p][1], name=f'd1[{p}]', norm_rtol=3e-2)
        assert_grad_close(D2.grad[p], want[p][2], name=f'd2[{p}]', norm_rtol=3e-2)


@pytest.mark.parametrize('variant,K,C,P', [('mast3r', 512, 768, 3), ('vggt', 300, 1024, 3), ('mast3r', 77, 64, 2),
                                           ('vggt', 130, 128, 2), ('mast3r', 257, 96, 2), ('vggt', 5, 64, 1)])
def test_smooth_ap_fused_kernel_forced(monkeypatch, variant, K, C, P):
    """The one-kernel path (similarity tile in TMEM, cluster-shared B tile) is chosen by a wave-fill heuristic that small
    batches never meet: force it (GD3_AP_FUSED=1) at the BASELINE sizes and at ragged K (1, 2, 3 and 4 CTAs per
    cluster, K not a multiple of 16 / 32) against the CPU oracle, and check that it equals the unfused path."""
    from gd3 import ops
    pairs = [make_pair(7600 + 10 * p, K, C) for p in range(P)]
    want = []
    for d1, d2, p1, p2 in pairs:
        a = d1.clone().requires_grad_(True)
        b = d2.clone().requires_grad_(True)
        loss = bodies.smooth_ap(a, b, p1, p2, variant)
        ga, gb = torch.autograd.grad(loss, [a, b])
        want.append((float(loss), ga, gb))
    D1, D2, P1, P2 = [torch.stack([q[k] for q in pairs]).cuda() for k in range(4)]
    got = {}
    for path in ('GD3_AP_FUSED', 'GD3_AP_UNFUSED'):
        monkeypatch.delenv('GD3_AP_FUSED', raising=False)
        monkeypatch.delenv('GD3_AP_UNFUSED', raising=False)
        monkeypatch.setenv(path, '1')
        x = D1.clone().requires_grad_(True)
        y = D2.clone().requires_grad_(True)
        loss = ops.smooth_ap(x, y, P1, P2, variant=variant)
        loss.sum().backward()
        got[path] = (loss.detach(), x.grad, y.grad)
        for p in range(P):
            assert rel_err(loss[p].item(), want[p][0]) <= 1e-3, (path, p, loss[p].item(), want[p][0])
            assert_grad_close(x.grad[p], want[p][1], name=f'{path} d1[{p}]', norm_rtol=3e-2)
            assert_grad_close(y.grad[p], want[p][2], name=f'{path} d2[{p}]', norm_rtol=3e-2)
    # forward-only call through the fused path (no d sim sweep)
    monkeypatch.delenv('GD3_AP_UNFUSED', raising=False)
    monkeypatch.setenv('GD3_AP_FUSED', '1')
    with torch.no_grad():
        l0 = ops.smooth_ap(D1, D2, P1, P2, variant=variant)
    assert torch.allclose(l0, got['GD3_AP_FUSED'][0], rtol=1e-6, atol=1e-7)
    assert torch.allclose(got['GD3_AP_FUSED'][0], got['GD3_AP_UNFUSED'][0], rtol=2e-4, atol=1e-6)


def test_smooth_ap_me_joint_mean_over_the_batch():
    """ME baseline with B > 1 (src/finetune_timm_me.py:199-217): one mean over the positives of all pairs.  Pairs get
    different numbers of positives (one gets none), so the per-pair mean ('me') and the joint mean differ."""
    from gd3 import ops
    K, C, B = 96, 64, 3
    pairs = [make_pair(7300 + 10 * p, K, C) for p in range(B)]
    D1, D2, P1, P2 = [torch.stack([q[k] for q in pairs]) for k in range(4)]
    P2 = P1 + 0.2                                   # nothing is a positive ...
    P2[0, :40] = P1[0, :40] + 1e-3                  # ... except 40 keypoints of pair 0 and 7 of pair 2
    P2[2, 5:12] = P1[2, 5:12] - 1e-3
    a = D1.clone().requires_grad_(True)
    b = D2.clone().requires_grad_(True)
    want = bodies.smooth_ap_me_batch(a, b, P1, P2)
    ga, gb = torch.autograd.grad(want, [a, b])
    x = D1.cuda().requires_grad_(True)
    y = D2.cuda().requires_grad_(True)
    loss = ops.smooth_ap(x, y, P1.cuda(), P2.cuda(), variant='me_joint')
    assert loss.shape == (B,) and float(loss[1]) == 0.0
    assert rel_err(loss.sum().item(), float(want)) <= 1e-3, (loss.tolist(), float(want))
    loss.sum().backward()
    assert_grad_close(x.grad, ga, name='d1', norm_rtol=3e-2)
    assert_grad_close(y.grad, gb, name='d2', norm_rtol=3e-2)
    assert float(x.grad[1].abs().max()) == 0.0
    per_pair = ops.smooth_ap(D1.cuda(), D2.cuda(), P1.cuda(), P2.cuda(), variant='me')
    assert math.isnan(per_pair[1].item()) and abs(per_pair[0].item() * 40 + per_pair[2].item() * 7 - 47 * float(want)) < 1e-3 * 47


def test_smooth_ap_edge_cases():
    from gd3 import ops
    # K = 0: the callers' early-out semantics -> loss 0, no kernel
    z = ops.smooth_ap(torch.zeros(2, 0, 16, device='cuda'), torch.zeros(2, 0, 16, device='cuda'),
                      torch.zeros(2, 0, 3, device='cuda'), torch.zeros(2, 0, 3, device='cuda'))
    assert z.shape == (2,) and float(z.abs().max()) == 0.0
    # 'me' without any positive: mean over an empty set is NaN in the reference
    d = torch.nn.functional.normalize(torch.randn(1, 8, 16, device='cuda'), dim=-1)
    p1 = torch.zeros(1, 8, 3, device='cuda')
    p2 = torch.ones(1, 8, 3, device='cuda')
    out = ops.smooth_ap(d, d, p1, p2, variant='me')
    assert math.isnan(out[0].item())
    with pytest.raises(ValueError):
        ops.smooth_ap(d, d, p1, p2, variant='nope')
    # forward only
    with torch.no_grad():
        l0 = ops.smooth_ap(d, d, p1, p1 + 0.001, variant='vggt')
    dd = d.clone().requires_grad_(True)
    l1 = ops.smooth_ap(dd, d, p1, p1 + 0.001, variant='vggt')
    assert abs(l0[0].item() - l1[0].item()) < 1e-6


@pytest.mark.parametrize('mode', ['all', 'proper', 'dual'])
def test_infonce_golden_and_gradients(golden, mode):
    """Upstream MASt3R InfoNCE: forward against the live reference's values, gradients against the oracle."""
    from gd3 import ops
    from oracle import losses as olosses
    g = golden('helpers.npz')
    d1, d2, vm = T(g['infonce/d1']), T(g['infonce/d2']), T(g['infonce/valid'])
    a = d1.cuda().requires_grad_(True)
    b = d2.cuda().requires_grad_(True)
    loss, rows = ops.infonce(a, b, vm.cuda(), mode=mode)
    assert rel_err(loss.item(), g[f'infonce/{mode}']) <= 1e-4
    loss.backward()
    ra = d1.clone().requires_grad_(True)
    rb = d2.clone().requires_grad_(True)
    olosses.infonce(ra, rb, vm, mode=mode).backward()
    assert_grad_close(a.grad, ra.grad, name='d1', norm_rtol=3e-2)
    assert_grad_close(b.grad, rb.grad, name='d2', norm_rtol=3e-2)
    # larger, all valid, descriptor size of MASt3R (24) and of the student (768)
    for K, D in ((300, 24), (512, 768)):
        gen = synth._gen(K + D)
        x = torch.nn.functional.normalize(torch.randn(2, K, D, generator=gen), dim=-1)
        y = torch.nn.functional.normalize(x + 0.2 * torch.randn(2, K, D, generator=gen), dim=-1)
        want = olosses.infonce(x, y, None, mode=mode)
        got, _ = ops.infonce(x.cuda(), y.cuda(), None, mode=mode)
        assert rel_err(got.item(), float(want)) <= 1e-3, (K, D, got.item(), float(want))
