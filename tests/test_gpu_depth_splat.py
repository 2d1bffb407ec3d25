"""gd3_point_cloud_to_depth against the live reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

from oracle import functions as ofn
from oracle import synth

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize('name', ['scene', 'sparse', 'none', 'halfpix'])
def test_point_cloud_to_depth_golden(golden, name):
    from gd3.compat import functions as cfn
    g = golden('depth_splat.npz')
    w, h = (int(v) for v in g[f'{name}/wh'])
    out = cfn.point_cloud_to_depth(T(g[f'{name}/pts']).cuda(), T(g[f'{name}/K']).cuda(), w, h, 'cuda')
    ref = g[f'{name}/depth']
    assert out.shape == ref.shape and out.dtype == torch.float32
    got = out.cpu().numpy()
    assert ((got > 0) == (ref > 0)).all()                 # every point lands on the reference's pixel
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=0)


def test_point_cloud_to_depth_batched_full_size():
    """Two 384 x 512 point maps with their own intrinsics in one call; each equals the oracle on its own."""
    from gd3 import _lib
    h, w = 384, 512
    Ks = torch.tensor([[[420.0, 0, 255.5], [0, 415.0, 191.5], [0, 0, 1]],
                       [[380.0, 0, 250.0], [0, 390.0, 200.0], [0, 0, 1]]])
    pts = torch.stack([synth.point_map(81 + b, h, w, float(Ks[b, 0, 0]), float(Ks[b, 1, 1]), float(Ks[b, 0, 2]),
                                       float(Ks[b, 1, 2]), oversample=1.0) for b in range(2)])
    out = _lib.point_cloud_to_depth(pts.cuda(), Ks.cuda(), w, h).cpu()
    again = _lib.point_cloud_to_depth(pts.cuda(), Ks.cuda(), w, h).cpu()
    assert torch.equal(out, again)                        # the double-precision sums make the splat order-independent
    for b in range(2):
        ref = ofn.point_cloud_to_depth(pts[b], Ks[b], w, h)[0, 0]
        assert ((out[b] > 0) == (ref > 0)).all()
        np.testing.assert_allclose(out[b].numpy(), ref.numpy(), rtol=1e-6, atol=0)
        assert (out[b] > 0).float().mean() > 0.3
    shared = _lib.point_cloud_to_depth(pts.cuda(), Ks[0].cuda(), w, h).cpu()
    assert torch.equal(shared[0], out[0])


def test_point_cloud_to_depth_empty():
    from gd3 import _lib
    K = torch.eye(3).cuda()
    out = _lib.point_cloud_to_depth(torch.zeros(1, 0, 3).cuda(), K, 8, 4)
    assert out.shape == (1, 4, 8) and float(out.abs().sum()) == 0.0
