"""gd3_vggt_attn_accumulate (VGGT half) against the live reference's golden maps and the oracle."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import teacher as oracle_teacher

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def _bf16_ulps(a, b):
    ia = (np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) >> 16).astype(np.int64)
    ib = (np.ascontiguousarray(b, dtype=np.float32).view(np.uint32) >> 16).astype(np.int64)
    return np.abs(ia - ib)


@pytest.mark.parametrize('name', ['small', 'ragged_t3', 'blocks'])
def test_vggt_attn_golden_fp32(golden, name):
    """fp32 semantics (round_bf16 = 0): head means per block and the block mean equal the live reference's."""
    from gd3 import _lib
    g = golden('vggt_attn.npz')
    B, heads, n, L = (int(v) for v in g[f'{name}/meta'])
    temp, scale = float(g[f'{name}/temperature']), float(g[f'{name}/scale'])
    maps = None
    for blk in range(L):
        q, k = T(g[f'{name}/q{blk}']).cuda(), T(g[f'{name}/k{blk}']).cuda()
        qs = (q * scale).bfloat16()                      # scale = 1 / 8: exact
        assert torch.equal(qs.float(), q * scale)
        a12, a21 = _lib.vggt_attn_accumulate(qs, k.bfloat16(), temp, 1.0, round_bf16=False)
        ref = T(g[f'{name}/attn_fp32_{blk}']).mean(dim=1)            # (2 B, n, n)
        np.testing.assert_allclose(a12.cpu().numpy(), ref[:B].numpy(), rtol=2e-5, atol=1e-8)
        np.testing.assert_allclose(a21.cpu().numpy(), ref[B:].numpy(), rtol=2e-5, atol=1e-8)
        maps = _lib.vggt_attn_accumulate(qs, k.bfloat16(), temp, 1.0 / L, out=maps, round_bf16=False) if maps is not None \
            else _lib.vggt_attn_accumulate(qs, k.bfloat16(), temp, 1.0 / L, round_bf16=False)
    np.testing.assert_allclose(maps[0].cpu().numpy(), g[f'{name}/cost_1'], rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(maps[1].cpu().numpy(), g[f'{name}/cost_2'], rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize('name', ['small', 'ragged_t3'])
def test_vggt_attn_golden_bf16_scores(golden, name):
    """bf16-autocast semantics, one head at a time: the maps, rounded to bf16, are the live reference's bf16 output
    (its softmax is evaluated in fp32 on the bf16 scores and rounded on output) up to rare one-step differences."""
    from gd3 import _lib
    g = golden('vggt_attn.npz')
    B, heads, n, _ = (int(v) for v in g[f'{name}/meta'])
    temp, scale = float(g[f'{name}/temperature']), float(g[f'{name}/scale'])
    q, k = T(g[f'{name}/q0']).cuda(), T(g[f'{name}/k0']).cuda()
    ref = g[f'{name}/attn_bf16_0']                                   # (2 B, heads, n, n)
    for h in range(heads):
        qs = (q[:, h:h + 1] * scale).bfloat16().contiguous()
        a12, a21 = _lib.vggt_attn_accumulate(qs, k[:, h:h + 1].bfloat16().contiguous(), temp, 1.0, round_bf16=True)
        got = torch.cat([a12, a21]).bfloat16().float().cpu().numpy()
        ulps = _bf16_ulps(got, ref[:, h])
        assert ulps.max() <= 1 and (ulps > 0).mean() < 0.01, (h, ulps.max(), (ulps > 0).mean())


@pytest.mark.parametrize('n,heads,B', [(925, 16, 1), (1369, 4, 2)])
def test_vggt_cost_volumes_full_size(n, heads, B):
    """VGGT's real shape (25 x 37 patches, 16 heads, head_dim 64) through the accumulator class, 2 blocks."""
    from gd3.compat import teacher
    vols = teacher.VggtCostVolumes(num_blocks=2, temperature=1.0)
    maps = []
    for blk in range(2):
        q, k = synth.vggt_qk(70 + blk, B, heads, n)
        vols.add_block(q.cuda(), k.cuda(), 0.125)
        maps.append(oracle_teacher.vggt_block_attention(q, k, 0.125, 1.0))
    c1, c2 = vols.result()
    r1, r2 = oracle_teacher.vggt_cost_volumes(maps)
    for got, ref in ((c1.cpu(), r1), (c2.cpu(), r2)):
        assert got.shape == ref.shape == (B, n, n)
        # a score that sits on a bf16 rounding boundary may round differently (fp32 summation order): one such flip
        # moves one probability by <= 2^-8 relative * |score|; bounded elementwise, negligible on average
        assert torch.allclose(got, ref, rtol=0.05, atol=1e-6)
        assert float((got - ref).abs().sum() / ref.abs().sum()) < 1e-4
        assert torch.allclose(got.sum(-1), torch.ones(B, n), atol=1e-5)


def test_vggt_attn_rejects_bad_inputs():
    from gd3 import _lib
    q = torch.zeros(1, 2, 21, 64, dtype=torch.bfloat16, device='cuda')
    with pytest.raises(ValueError):
        _lib.vggt_attn_accumulate(q, q)                  # odd token count
    q = torch.zeros(1, 2, 8, 64, dtype=torch.bfloat16, device='cuda')
    with pytest.raises(ValueError):
        _lib.vggt_attn_accumulate(q, q)                  # nothing left after the 5 special tokens
    with pytest.raises(ValueError):
        _lib.vggt_attn_accumulate(q.float(), q.float())
