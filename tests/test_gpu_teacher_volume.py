"""GPU parity: fused MASt3R teacher cost-volume post-processing (dust3r/dust3r/model.py:346-363)."""
import pytest
import torch

from oracle import teacher as oracle_teacher

pytestmark = pytest.mark.gpu


def _maps(seed, L, B, H, N):
    g = torch.Generator().manual_seed(seed)
    tgt = [3.0 * torch.randn(B, H, N, N, generator=g) for _ in range(L)]
    src = [3.0 * torch.randn(B, H, N, N, generator=g) for _ in range(L)]
    return tgt, src


@pytest.mark.parametrize('L,B,H,N', [(2, 1, 3, 64), (3, 2, 4, 100), (12, 1, 12, 256), (2, 1, 2, 37 * 37), (1, 3, 1, 33)])
@pytest.mark.parametrize('reciprocity', [True, False])
def test_teacher_volume_matches_oracle(L, B, H, N, reciprocity):
    from gd3.compat import teacher
    tgt, src = _maps(L * 100 + N, L, B, H, N)
    ref = oracle_teacher.teacher_volume([t.clone() for t in tgt], [s.clone() for s in src], 3.0, reciprocity)
    got = teacher.teacher_volume([t.cuda() for t in tgt], [s.cuda() for s in src], 3.0, reciprocity).cpu()
    assert got.shape == ref.shape == (B, N, N)
    # fp32 throughout; only the association of the head / layer sums differs
    assert torch.allclose(got, ref, rtol=2e-5, atol=1e-7), (got - ref).abs().max()
    # column 0 holds the mean of the layer minima, identical for every row and batch entry
    assert torch.equal(got[:, :, 0], got[:1, :1, 0].expand(B, N))
    if reciprocity:
        rest = got[:, :, 1:].sum(-1)
        assert (rest < 1.0).all() and (rest > 0.5).all()     # rows of a softmax minus one column


def test_teacher_volume_feeds_cost_kl():
    """The fused producer's output is a valid teacher volume for the KL loss (rows re-normalised there)."""
    from gd3 import ops
    from gd3.compat import teacher
    L, H, N, C = 4, 4, 256, 64
    tgt, src = _maps(9, L, 1, H, N)
    t12 = teacher.teacher_volume([t.cuda() for t in tgt], [s.cuda() for s in src], 3.0, True)
    t21 = teacher.teacher_volume([s.cuda() for s in src], [t.cuda() for t in tgt], 3.0, True)
    g = torch.Generator().manual_seed(1)
    f1 = torch.randn(1, N, C, generator=g).cuda().to(torch.bfloat16)
    f2 = torch.randn(1, N, C, generator=g).cuda().to(torch.bfloat16)
    m = torch.ones(1, N, dtype=torch.bool).cuda()
    loss = ops.cost_volume_kl(f1, f2, t12, t21, m, m, variant='mast3r')
    assert torch.isfinite(loss).all() and (loss > 0).all()
    # the same producer handing the volume over in the packed form (fp16 * 1024 + row statistics): same loss
    p12 = teacher.teacher_volume([t.cuda() for t in tgt], [s.cuda() for s in src], 3.0, True, packed=True)
    p21 = teacher.teacher_volume([s.cuda() for s in src], [t.cuda() for t in tgt], 3.0, True, packed=True)
    assert p12[0].dtype == torch.float16 and p12[1].shape == (1, 3, N)
    assert torch.allclose(p12[0].float() / 1024.0, t12, rtol=1e-3, atol=1e-7)
    assert torch.allclose(p12[1][:, 0], t12.sum(-1), rtol=1e-5)
    lp = ops.cost_volume_kl(f1, f2, p12, p21, m, m, variant='mast3r')
    assert abs(lp.item() - loss.item()) <= 1e-3 * abs(loss.item())


def test_vggt_plain_mean_matches_oracle():
    from gd3.compat import teacher
    g = torch.Generator().manual_seed(4)
    maps = [torch.softmax(2 * torch.randn(2, 5, 120, 120, generator=g), -1) for _ in range(6)]
    r1, r2 = oracle_teacher.vggt_cost_volumes(maps)
    c1, c2 = teacher.vggt_cost_volumes([m.cuda() for m in maps])
    assert torch.allclose(c1.cpu(), r1, rtol=2e-5, atol=1e-8) and torch.allclose(c2.cpu(), r2, rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize('name', ['recip', 'recip_t1', 'plain'])
def test_teacher_volume_live_reference_golden(golden, name):
    """The fused kernel against ``tgt_attn_map`` produced by the live MASt3R teacher class on its own logits."""
    import numpy as np
    from gd3.compat import teacher
    g = golden('teacher_volume.npz')
    tgt = [t.cuda() for t in torch.from_numpy(g[f'{name}/tgt'])]
    src = [t.cuda() for t in torch.from_numpy(g[f'{name}/src'])]
    got = teacher.teacher_volume(tgt, src, float(g[f'{name}/temperature']), bool(g[f'{name}/reciprocity']))
    np.testing.assert_allclose(got.cpu().numpy(), g[f'{name}/tgt_attn_map'], rtol=2e-5, atol=1e-7)
