"""GPU parity of the fused relative-depth losses (ranking / hinge / cross-view L1) through the C ABI
against golden vectors of the live reference and the CPU oracle."""
import pytest
import torch

from oracle import bodies, losses as olosses, synth
from helpers import assert_grad_close, rel_err
from test_oracle_golden import RANK_CASES, load_head_case

pytestmark = pytest.mark.gpu
T = torch.as_tensor
PNAMES = ['W1', 'b1', 'gamma', 'beta', 'w2', 'b2']


def cuda_head(head):
    import copy
    return copy.deepcopy(head).cuda()


def head_grads(head):
    fl = head.fusion_layer
    ps = [fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight, fl[3].bias]
    return [torch.zeros_like(p) if p.grad is None else p.grad for p in ps]


@pytest.mark.parametrize('case', RANK_CASES)
def test_ranking_and_hinge_golden(golden, case):
    """Drop-in signatures of utils/losses.py against the live reference's values and gradients."""
    from gd3.compat import losses
    g = golden('ranking.npz')
    head, kf1, kf2, kd1, kd2 = load_head_case(g, case)
    for tag, fn, kw in (('rank', losses.pairwise_logistic_ranking_loss, dict(depth_threshold=0.05)),
                        ('hinge', losses.intra_depth_loss, {})):
        h = cuda_head(head)
        x = kf1.detach().cuda().requires_grad_(True)
        if tag == 'rank':
            # golden 'rank' = (ranking(set 1) + ranking(set 2)) / 2
            y = kf2.detach().cuda().requires_grad_(True)
            val = (fn(h, x, kd1.cuda(), **kw) + fn(h, y, kd2.cuda(), **kw)) / 2
        else:
            y = None
            val = fn(h, x, kd1.cuda(), **kw)
        want = float(g[f'{case}/{tag}/loss'])
        assert abs(val.item() - want) <= 1e-3 * abs(want) + 1e-7, (tag, val.item(), want)
        if want == 0.0:
            continue
        val.backward()
        assert_grad_close(x.grad, T(g[f'{case}/{tag}/grad_kf1']), name=f'{tag}/kf1', norm_rtol=3e-2)
        if y is not None:
            assert_grad_close(y.grad, T(g[f'{case}/{tag}/grad_kf2']), name=f'{tag}/kf2', norm_rtol=3e-2)
        for n_, got in zip(PNAMES, head_grads(h)):
            assert_grad_close(got, T(g[f'{case}/{tag}/grad_{n_}']), name=f'{tag}/{n_}', norm_rtol=3e-2)


@pytest.mark.parametrize('case', ['small', 'cfg1', 'notanh'])
def test_fused_depth_losses_golden(golden, case):
    """One fused call = calculate_depth_loss: ranking on both views + cross-view L1, shared head gradients."""
    from gd3 import ops
    g = golden('ranking.npz')
    head, kf1, kf2, kd1, kd2 = load_head_case(g, case)
    h = cuda_head(head)
    feats = torch.cat([kf1.detach(), kf2.detach()]).cuda().requires_grad_(True)     # sets (view 1, view 2)
    depths = torch.cat([kd1, kd2]).cuda()
    w_rank = torch.tensor([0.5, 0.5], device='cuda')
    w_l1 = torch.tensor([1.0], device='cuda')
    total, lr, l1 = ops.depth_head_loss(h, feats, depths, mode='logistic', depth_threshold=0.05,
                                        w_rank=w_rank, w_l1=w_l1)
    want_rank, want_l1 = float(g[f'{case}/rank/loss']), float(g[f'{case}/l1/loss'])
    assert rel_err(0.5 * (lr[0] + lr[1]).item(), want_rank) <= 1e-3
    assert rel_err(l1[0].item(), want_l1) <= 1e-3
    assert rel_err(total.item(), want_rank + want_l1) <= 1e-3
    total.backward()
    want_f = torch.cat([T(g[f'{case}/rank/grad_kf1']) + T(g[f'{case}/l1/grad_kf1']),
                        T(g[f'{case}/rank/grad_kf2']) + T(g[f'{case}/l1/grad_kf2'])])
    assert_grad_close(feats.grad, want_f, name='feats', norm_rtol=3e-2)
    for n_, got in zip(PNAMES, head_grads(h)):
        want = T(g[f'{case}/rank/grad_{n_}']) + T(g[f'{case}/l1/grad_{n_}'])
        assert_grad_close(got, want, name=n_, norm_rtol=3e-2)


@pytest.mark.parametrize('K,D', [(512, 768), (300, 1024)])
def test_depth_losses_full_size(K, D):
    """BASELINE.json sizes (cfg2 K=512,D=768; cfg4 K=300,D=1024), 2 pairs per call, vs the CPU oracle."""
    from gd3 import ops
    P = 2
    head = olosses.DepthHead(D)
    synth.load_head(head, synth.head_params(81, D))
    gen = synth._gen(82)
    feats = 0.5 * torch.randn(2 * P, K, D, generator=gen)
    depths = torch.stack([synth.depths(90 + s, K) for s in range(2 * P)])
    ref_f = feats.clone().requires_grad_(True)
    tot = 0.0
    want_rank, want_l1 = [], []
    for p in range(P):
        l1, rk = bodies.depth_losses(head, ref_f[2 * p:2 * p + 1], ref_f[2 * p + 1:2 * p + 2],
                                     depths[2 * p:2 * p + 1], depths[2 * p + 1:2 * p + 2])
        tot = tot + rk + 0.7 * l1
        want_rank.append(float(rk))
        want_l1.append(float(l1))
    tot.backward()
    want_p = head_grads(head)
    h = cuda_head(head)
    for prm in h.parameters():
        prm.grad = None
    x = feats.cuda().requires_grad_(True)
    total, lr, l1 = ops.depth_head_loss(h, x, depths.cuda(), w_rank=torch.full((2 * P,), 0.5, device='cuda'),
                                        w_l1=torch.full((P,), 0.7, device='cuda'))
    total.backward()
    for p in range(P):
        assert rel_err(0.5 * (lr[2 * p] + lr[2 * p + 1]).item(), want_rank[p]) <= 1e-3
        assert rel_err(l1[p].item(), want_l1[p]) <= 1e-3
    assert rel_err(total.item(), float(tot)) <= 1e-3
    assert_grad_close(x.grad, ref_f.grad, name='feats', norm_rtol=3e-2)
    for n_, got, want in zip(PNAMES, head_grads(h), want_p):
        assert_grad_close(got, want, name=n_, norm_rtol=3e-2)


def test_depth_losses_batch_split_invariance():
    """Size-independent property at the strong-scaling shard size of cfg4 (8 pairs per GPU, K = 300, D = 1024): the
    per-set losses and the feature gradients of a batch equal those of its halves evaluated separately.  The pair
    kernel picks a different b tile for every (sets, K) -- 7 rows per warp for 16 sets, another for 8 -- so this also
    pins the launch-shape model of rank_b_per_warp."""
    from gd3 import ops
    P, K, D = 8, 300, 1024
    head = olosses.DepthHead(D)
    synth.load_head(head, synth.head_params(83, D))
    h = cuda_head(head)
    gen = synth._gen(84)
    feats = (0.5 * torch.randn(2 * P, K, D, generator=gen)).cuda()
    depths = torch.stack([synth.depths(120 + s, K) for s in range(2 * P)]).cuda()

    def run(lo, hi):
        x = feats[2 * lo:2 * hi].clone().requires_grad_(True)
        n = hi - lo
        total, lr, l1 = ops.depth_head_loss(h, x, depths[2 * lo:2 * hi], w_rank=torch.full((2 * n,), 0.5, device='cuda'),
                                            w_l1=torch.full((n,), 0.7, device='cuda'))
        total.backward()
        return lr.detach(), l1.detach(), x.grad

    lr, l1, g = run(0, P)
    lr_a, l1_a, g_a = run(0, P // 2)
    lr_b, l1_b, g_b = run(P // 2, P)
    assert torch.allclose(lr, torch.cat([lr_a, lr_b]), rtol=1e-5, atol=1e-7)
    assert torch.allclose(l1, torch.cat([l1_a, l1_b]), rtol=1e-5, atol=1e-7)
    g_split = torch.cat([g_a, g_b])
    assert_grad_close(g, g_split, name='feats (whole batch vs halves)', norm_rtol=1e-3)
    assert float((g - g_split).abs().max()) <= 1e-3 * float(g.abs().max())


def test_depth_losses_edge_cases():
    from gd3 import ops
    from gd3.compat import losses
    head = olosses.DepthHead(32).cuda()
    f = torch.randn(1, 9, 32, device='cuda', requires_grad=True)
    d = torch.ones(1, 9, device='cuda')
    # no valid pair: the reference returns a constant 0
    out = losses.pairwise_logistic_ranking_loss(head, f, d, depth_threshold=0.05)
    assert out.item() == 0.0
    out.backward()
    assert float(f.grad.abs().max()) == 0.0
    # K = 0 / forward only
    z, lr, _ = ops.depth_head_loss(head, torch.zeros(2, 0, 32, device='cuda'), torch.zeros(2, 0, device='cuda'))
    assert z.item() == 0.0 and lr.shape == (2,)
    with torch.no_grad():
        a = losses.intra_depth_loss(head, f, torch.rand(1, 9, device='cuda') * 3)
    assert a.item() >= 0.0
    with pytest.raises(ValueError):
        ops.depth_head_loss(olosses.DepthHead(32, hidden_dim=64).cuda(), f, d)
    # ragged K (not a multiple of the 128 tile or of 8) and joint mean over B = 3 sets
    K = 77
    feats = torch.randn(3, K, 32)
    depths = torch.rand(3, K) * 4
    hc = olosses.DepthHead(32)
    want = olosses.pairwise_logistic_ranking_loss(hc, feats, depths, 0.05)
    got = losses.pairwise_logistic_ranking_loss(cuda_head(hc), feats.cuda(), depths.cuda(), 0.05)
    assert rel_err(got.item(), float(want)) <= 1e-3


@pytest.mark.parametrize('spread', [1.0, 0.1, 0.01, 0.001])
def test_ranking_correlated_keypoint_features(spread):
    """Keypoint features that share a large common component (real ViT tokens of one image do): the LayerNorm
    statistics of the pair differences come from a Gram matrix, which must not lose the small differences."""
    from gd3.compat import losses
    torch.manual_seed(11)
    K, D = 96, 64
    head = olosses.DepthHead(D)
    with torch.no_grad():
        head.fusion_layer[0].bias.mul_(0.05)          # small bias: the pair statistics are dominated by f_j - f_i
    base = torch.randn(1, 1, D) * 3.0
    feats = base + spread * torch.randn(1, K, D)
    feats[0, 5] = feats[0, 4] + 1e-4 * spread * torch.randn(D)        # a near-duplicate pair on top
    depths = torch.rand(1, K) * 4.5 + 0.5
    ref_in = feats.clone().requires_grad_(True)
    want = olosses.pairwise_logistic_ranking_loss(head, ref_in, depths, depth_threshold=0.05)
    want.backward()
    h = cuda_head(head)
    x = feats.cuda().requires_grad_(True)
    got = losses.pairwise_logistic_ranking_loss(h, x, depths.cuda(), depth_threshold=0.05)
    got.backward()
    assert rel_err(got.item(), want.item()) <= 1e-3, (spread, got.item(), want.item())
    assert_grad_close(x.grad, ref_in.grad, name=f'feats/spread={spread}', norm_rtol=3e-2)


def test_ranking_exact_duplicate_keypoints():
    """Many keypoints with EXACTLY the same feature (several keypoints on one token): their pair differences are the
    bias alone, the Gram form of the LayerNorm statistics cancels completely, and every such pair goes through the
    direct recomputation (rank_fix_rstd; rows with many flags, K not a multiple of 32)."""
    from gd3.compat import losses
    torch.manual_seed(13)
    K, D = 77, 64
    head = olosses.DepthHead(D)
    feats = torch.randn(2, K, D)
    feats[0, 10:30] = feats[0, 10]                 # 20 identical rows in set 0
    feats[1, 0::7] = feats[1, 0]                   # every 7th row identical in set 1
    depths = torch.rand(2, K) * 4.5 + 0.5
    ref_in = feats.clone().requires_grad_(True)
    want = olosses.pairwise_logistic_ranking_loss(head, ref_in, depths, depth_threshold=0.05)
    want.backward()
    h = cuda_head(head)
    x = feats.cuda().requires_grad_(True)
    got = losses.pairwise_logistic_ranking_loss(h, x, depths.cuda(), depth_threshold=0.05)
    got.backward()
    assert torch.isfinite(x.grad).all()
    assert rel_err(got.item(), want.item()) <= 1e-3, (got.item(), want.item())
    assert_grad_close(x.grad, ref_in.grad, name='feats with duplicates', norm_rtol=3e-2)
    for i, (p_ref, p_got) in enumerate(zip(head_grads(head), head_grads(h))):
        assert_grad_close(p_got.cpu().reshape(-1), p_ref.reshape(-1), name=f'head[{i}]', norm_rtol=3e-2)


def test_depth_head_loss_rejects_negative_threshold():
    """thr >= 0 is part of the contract (a pair of equal depths is never valid): the library says so instead of
    silently counting the diagonal."""
    from gd3 import _lib
    from gd3.compat import losses
    head = cuda_head(olosses.DepthHead(32))
    with pytest.raises((ValueError, _lib.Gd3Error), match='thr must be >= 0'):
        losses.pairwise_logistic_ranking_loss(head, torch.randn(1, 8, 32).cuda(), torch.rand(1, 8).cuda(), -0.1)
