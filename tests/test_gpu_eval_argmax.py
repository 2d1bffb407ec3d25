"""GPU parity: keypoint -> arg-max pixel transfer (src/evaluate_timm.py:532-547) through the C ABI."""
import pytest
import torch

from oracle import evaluate as oracle_eval

pytestmark = pytest.mark.gpu


def _case(seed, C, ph, K, smooth=True):
    g = torch.Generator().manual_seed(seed)
    d2 = torch.randn(1, C, ph, ph, generator=g)
    if smooth:   # neighbouring patches correlate like real ViT features
        d2 = torch.nn.functional.avg_pool2d(d2, 3, stride=1, padding=1)
    kd = torch.nn.functional.normalize(torch.randn(1, C, K, generator=g), dim=1)
    return kd, d2


@pytest.mark.parametrize('img,patch,stride,C,K', [(224, 14, 14, 384, 20), (224, 14, 7, 96, 5), (448, 14, 14, 768, 30),
                                                   (240, 16, 16, 64, 1), (224, 14, 14, 33, 17)])
def test_semantic_argmax_matches_oracle(img, patch, stride, C, K):
    from gd3.compat import evaluate
    ph = 1 + (img - patch) // stride
    kd, d2 = _case(img + C + K, C, ph, K)
    ref_idx, sim = oracle_eval.semantic_argmax(kd, d2, img, patch, stride)
    idx, xy = evaluate.semantic_argmax(kd.cuda(), d2.cuda(), img, patch, stride)
    idx = idx.cpu()
    assert idx.dtype == torch.int64 and idx.shape == (K,)
    assert (xy.cpu()[:, 0] == idx % img).all() and (xy.cpu()[:, 1] == idx // img).all()
    # the upsampled map is never built here, so sums are associated differently: a mismatch must be a near-tie
    best = sim.max(dim=1).values
    at = sim[torch.arange(K), idx]
    assert ((best - at).abs() <= 2e-6 * best.abs().clamp_min(1.0)).all()
    assert (idx == ref_idx).float().mean() >= 0.8


def test_semantic_argmax_exact_and_ties():
    """Exactly representable data: bit-identical similarities, ties to the lowest pixel index."""
    from gd3 import _lib
    img, patch, stride = 56, 14, 14
    ph = 1 + (img - patch) // stride            # 4 patches; ds = 43
    C, K = 8, 6
    g = torch.Generator().manual_seed(3)
    d2 = torch.randint(-2, 3, (1, C, ph, ph), generator=g).float()
    kd = torch.randint(-2, 3, (1, C, K), generator=g).float()
    d2[0, :, 0, 0] = d2[0, :, 3, 3]             # equal corner patches -> equal padded borders
    ref_idx, sim = oracle_eval.semantic_argmax(kd, d2, img, patch, stride)
    idx, val = _lib.semantic_argmax(kd.cuda(), d2.cuda(), img, patch, stride, want_val=True)
    best = sim.max(dim=1).values
    assert torch.allclose(val.cpu(), best, rtol=1e-6, atol=1e-6)
    at = sim[torch.arange(K), idx.cpu()]
    assert torch.allclose(at, best, rtol=1e-6, atol=1e-6)
    # constant map: every pixel ties -> index 0
    d2c = torch.ones(1, C, ph, ph)
    idx = _lib.semantic_argmax(kd.cuda(), d2c.cuda(), img, patch, stride)
    assert (idx.cpu() == 0).all()
    # (K, C) layout of the keypoint descriptors gives the same answer
    i1 = _lib.semantic_argmax(kd.cuda(), d2.cuda(), img, patch, stride)
    i2 = _lib.semantic_argmax(kd[0].t().contiguous().cuda(), d2.cuda(), img, patch, stride)
    assert torch.equal(i1, i2)


def test_semantic_argmax_vs_live_semantic_transfer(golden):
    """The CUDA path (``interpolate_features`` + ``semantic_argmax``) against ``nn_idx`` recorded inside the reference's
    own ``semantic_transfer`` run live (``tests/golden/eval_argmax.npz``)."""
    import numpy as np
    from gd3.compat import evaluate
    from gd3.compat import functions as cfn
    g = golden('eval_argmax.npz')
    img, patch, stride, ph, C, K = (int(v) for v in g['meta'])
    d1 = torch.from_numpy(g['tokens1']).reshape(1, ph, ph, C).permute(0, 3, 1, 2).contiguous().cuda()
    d2 = torch.from_numpy(g['tokens2']).reshape(1, ph, ph, C).permute(0, 3, 1, 2).contiguous()
    kps = torch.from_numpy(g['kps1'])[None, :, :2].cuda()
    kd = cfn.interpolate_features(d1, kps, h=img, w=img, normalize=True)
    idx, xy = evaluate.semantic_argmax(kd, d2.cuda(), img, patch, stride)
    idx = idx.cpu().numpy()
    # a different index is acceptable only as a rounding-level near-tie of the similarity (checked on the oracle's map)
    _, sim = oracle_eval.semantic_argmax(kd.cpu(), d2, img, patch, stride)
    at = sim[torch.arange(K), torch.from_numpy(idx)].numpy()
    assert (np.abs(at - g['best']) <= 2e-6 * np.maximum(1.0, np.abs(g['best']))).all()
    assert (idx == g['nn_idx']).mean() >= 0.9
