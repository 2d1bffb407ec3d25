"""GPU parity of the volume-level helpers (get_masked_patch_cost, kl_divergence_map) through the C ABI:
forward and backward against the CPU oracle in fp64 / fp32 and against the live reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import functions as ofn
from oracle import losses as olosses
from helpers import assert_grad_close, rel_err

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _volume(seed, B, n1, n2, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(B, n1, n2, generator=g)


# row lengths: register-resident kernels (<= 1024, <= 2048, 16-byte rows), streaming vector kernel (> 2048), scalar (ragged)
@pytest.mark.parametrize('B,n1,n2', [(1, 64, 64), (2, 37, 50), (2, 1369, 1369), (1, 5, 3), (2, 40, 1368), (1, 8, 2052)])
@pytest.mark.parametrize('mode', ['rownorm', 'softmax', 'softmax_t', 'rownorm_m2', 'softmax_m2'])
def test_masked_patch_cost_forward_backward(B, n1, n2, mode):
    from gd3.compat import functions as fn
    g = torch.Generator().manual_seed(7 + n1)
    cost = _volume(11 + n2, B, n1, n2)
    if mode.startswith('rownorm'):
        cost = cost.abs() * 1e-2            # teacher-like: non-negative
    m1 = torch.rand(n1, generator=g) < 0.6
    m1[0] = False                          # at least one masked row
    m2 = (torch.rand(n2, generator=g) < 0.7) if mode.endswith('_m2') else None
    kw = dict(use_softmax=mode.startswith('softmax'), temperature=0.07 if mode == 'softmax_t' else 1.0)
    w = _volume(13, B, n1, n2)             # random cotangent
    # oracle in fp64
    xr = cost.double().requires_grad_(True)
    yr = ofn.get_masked_patch_cost(xr, m1, m2, **kw)
    (yr.double() * w.double()).sum().backward()
    x = cost.cuda().requires_grad_(True)
    y = fn.get_masked_patch_cost(x, m1.cuda(), None if m2 is None else m2.cuda(), **kw)
    assert y.dtype == torch.float32 and y.shape == cost.shape
    (y * w.cuda()).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().float().numpy(), rtol=2e-5, atol=1e-9)
    assert_grad_close(x.grad, xr.grad.float(), cos_min=0.99999, name=f'cost ({mode})', norm_rtol=1e-4)
    # masked rows / columns get exactly no gradient
    assert float(x.grad[:, ~m1.cuda()].abs().sum()) == 0.0
    if m2 is not None:
        assert float(x.grad[:, :, ~m2.cuda()].abs().sum()) == 0.0


def test_masked_patch_cost_edge_cases():
    from gd3.compat import functions as fn
    # a kept row whose sum is below eps: divided by eps, and the clamp passes no gradient through the sum
    cost = torch.zeros(1, 4, 8)
    cost[0, 1] = 1e-10
    cost[0, 2, 3] = 2.0
    m1 = torch.tensor([True, True, True, False])
    for kw in (dict(), dict(use_softmax=True, temperature=0.5)):
        xr = cost.clone().requires_grad_(True)
        yr = ofn.get_masked_patch_cost(xr, m1, None, **kw)
        yr.square().sum().backward()
        x = cost.cuda().requires_grad_(True)
        y = fn.get_masked_patch_cost(x, m1.cuda(), **kw)
        y.square().sum().backward()
        np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), rtol=1e-6, atol=0)
        np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-9)
    # all rows masked: zeros (row-normalised) / uniform rows (softmax)
    none = torch.zeros(4, dtype=torch.bool)
    assert float(fn.get_masked_patch_cost(cost.cuda(), none.cuda()).abs().max()) == 0.0
    u = fn.get_masked_patch_cost(cost.cuda(), none.cuda(), use_softmax=True)
    assert torch.allclose(u, torch.full_like(u, 1.0 / 8))
    with pytest.raises(ValueError):
        fn.get_masked_patch_cost(cost.cuda(), torch.ones(5, dtype=torch.bool).cuda())
    with pytest.raises(ValueError):
        fn.get_masked_patch_cost(cost, m1)           # CPU tensors: no fallback


@pytest.mark.parametrize('shape', [(1, 64, 64), (2, 37, 50), (3, 5, 3), (2, 1369, 1369), (4, 1024, 1024)])
def test_kl_divergence_map_forward_backward(shape):
    from gd3.compat import losses
    g = torch.Generator().manual_seed(sum(shape))
    t = torch.softmax(4.0 * torch.randn(*shape, generator=g), dim=-1)      # peaked: many entries below eps
    s = torch.softmax(torch.randn(*shape, generator=g), dim=-1)
    t[..., 0] = 0.0                                                       # exact zeros hit the clamp
    s[0, 0, :2] = 1e-9
    tr, sr = t.double().requires_grad_(True), s.double().requires_grad_(True)
    want = olosses.kl_divergence_map(tr, sr)
    want.backward()
    tc, sc = t.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    got = losses.kl_divergence_map(tc, sc)
    assert got.dim() == 0 and got.dtype == torch.float32
    (3.0 * got).backward()
    assert rel_err(got.item(), want.item()) <= 2e-6
    assert_grad_close(sc.grad, 3.0 * sr.grad.float(), cos_min=0.999999, name='student', norm_rtol=1e-5)
    assert_grad_close(tc.grad, 3.0 * tr.grad.float(), cos_min=0.999999, name='teacher', norm_rtol=1e-5)
    assert float((sc.grad - 3.0 * sr.grad.float().cuda()).abs().max()) <= 1e-5 * float(sr.grad.abs().max()) * 3.0
    # only the student needs a gradient (the training case): the teacher's is not computed
    sc2 = s.cuda().requires_grad_(True)
    losses.kl_divergence_map(t.cuda(), sc2).backward()
    assert_grad_close(sc2.grad, sr.grad.float(), cos_min=0.999999, name='student only', norm_rtol=1e-5)
    # deterministic: the row sums are added in a fixed order
    a = losses.kl_divergence_map(t.cuda(), s.cuda()).item()
    assert a == losses.kl_divergence_map(t.cuda(), s.cuda()).item()


def test_volume_helpers_golden(golden):
    """The live reference's outputs (tests/golden/helpers.npz, oracle/gen_golden.py)."""
    from gd3.compat import losses
    g = golden('helpers.npz')
    got = losses.kl_divergence_map(T(g['klmap/t']).cuda(), T(g['klmap/s']).cuda())
    assert rel_err(got.item(), float(g['klmap/out'])) < 1e-6
    with pytest.raises(ValueError):
        losses.kl_divergence_map(torch.rand(2, 4, 4).cuda(), torch.rand(2, 4, 5).cuda())


def test_minimal_drop_in_matches_fused_kl():
    """INTEGRATION.md section 2 vs section 3: the volume-level helpers chained the way calculate_cost_loss chains them
    (src/finetune_timm_mast3r.py:521-540) give the loss and feature gradients of the fused gd3_cost_kl."""
    from gd3 import ops
    from gd3.compat import functions as fn, losses
    P, N, C = 2, 256, 384
    g = torch.Generator().manual_seed(902)
    f1 = torch.randn(P, N, C, generator=g).cuda().requires_grad_(True)
    f2 = torch.randn(P, N, C, generator=g).cuda().requires_grad_(True)
    t12 = torch.softmax(3.0 * torch.randn(P, N, N, generator=g), -1).cuda()
    t21 = torch.softmax(3.0 * torch.randn(P, N, N, generator=g), -1).cuda()
    m1 = (torch.rand(P, N, generator=g) < 0.6).cuda()
    m2 = (torch.rand(P, N, generator=g) < 0.6).cuda()
    total = 0.0
    for p in range(P):
        a = torch.nn.functional.normalize(f1[p:p + 1], dim=-1)
        b = torch.nn.functional.normalize(f2[p:p + 1], dim=-1)
        z12 = torch.bmm(a, b.transpose(1, 2))
        z21 = torch.bmm(b, a.transpose(1, 2))
        kl12 = losses.kl_divergence_map(fn.get_masked_patch_cost(t12[p:p + 1], m1[p]),
                                        fn.get_masked_patch_cost(z12, m1[p], use_softmax=True))
        kl21 = losses.kl_divergence_map(fn.get_masked_patch_cost(t21[p:p + 1], m2[p]),
                                        fn.get_masked_patch_cost(z21, m2[p], use_softmax=True))
        total = total + (kl12 + kl21) / 2
    total.backward()
    g1, g2 = f1.grad.clone(), f2.grad.clone()
    f1.grad = f2.grad = None
    fused = ops.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant='mast3r')
    fused.sum().backward()
    assert rel_err(fused.sum().item(), total.item()) <= 1e-3
    assert_grad_close(f1.grad, g1, name='f1')
    assert_grad_close(f2.grad, g2, name='f2')
