"""Seeded synthetic inputs for the hot path.

Shared by the tests, ``bench.py`` and ``oracle/gen_golden.py`` so that the CUDA
path, the oracle and the live reference always see identical tensors.  Pure CPU
torch; no reference code involved.  Test infrastructure only.
"""

import math
from dataclasses import dataclass

import torch

# BASELINE.json configs (the loss configs; cfg3 is the fast_nn case)
CONFIGS = {
    'cfg1': dict(N=256, C=384, K=128, grid=(16, 16), P=1, variant='mast3r'),
    'cfg2': dict(N=1024, C=768, K=512, grid=(32, 32), P=32, variant='mast3r'),
    'cfg4': dict(N=1369, C=1024, K=300, grid=(37, 37), P=64, variant='vggt'),
}


def _gen(seed):
    g = torch.Generator(device='cpu')
    g.manual_seed(int(seed))
    return g


def bf16_round(x):
    """Round to bf16 and come back: the *same* values feed the CUDA path and the oracle."""
    return x.to(torch.bfloat16).to(torch.float32)


def features(seed, N, C):
    """Student patch features f1, f2 ~ N(0, 1), bf16-exact fp32 (N, C)."""
    g = _gen(seed)
    f1 = bf16_round(torch.randn(N, C, generator=g))
    f2 = bf16_round(torch.randn(N, C, generator=g))
    return f1, f2


def teacher_volume(seed, N, variant='mast3r', scale=4.0, peak=6.0, heads=16):
    """One direction of a teacher cost volume, (N, N) fp32, rows roughly stochastic.

    'mast3r': softmax of peaked logits, then column 0 overwritten with the global
    minimum (mimics ``dust3r/dust3r/model.py:352-354``), not renormalised.
    'vggt': mean over ``heads`` row-softmaxes of such logits
    (``vggt/layers/attention.py:73-84`` + head/block means).
    """
    g = _gen(seed)
    perm = torch.randperm(N, generator=g)
    rows = torch.arange(N)
    if variant == 'mast3r':
        logits = scale * torch.randn(N, N, generator=g)
        logits[rows, perm] += peak
        t = torch.softmax(logits, dim=-1)
        t[:, 0] = t.min()
        return t.contiguous()
    acc = torch.zeros(N, N)
    for _ in range(heads):
        logits = scale * torch.randn(N, N, generator=g)
        logits[rows, perm] += peak
        acc += torch.softmax(logits, dim=-1)
    return (acc / heads).contiguous()


def patch_mask(seed, N, p_keep=0.6, mode='bernoulli'):
    """(N,) bool patch mask; always >= 1 kept row in 'bernoulli' mode."""
    if mode == 'all':
        return torch.ones(N, dtype=torch.bool)
    if mode == 'none':
        return torch.zeros(N, dtype=torch.bool)
    g = _gen(seed)
    m = torch.rand(N, generator=g) < p_keep
    m[int(torch.randint(0, N, (1,), generator=g))] = True
    return m


def ap_token_maps(seed, N, C):
    """Correlated token maps for the Smooth-AP loss: sampled cosines land in ~[0.96, 1]."""
    g = _gen(seed)
    base = torch.randn(1, C, generator=g)
    g1 = base + 0.12 * torch.randn(N, C, generator=g)
    g2 = g1 + 0.03 * torch.randn(N, C, generator=g)
    return bf16_round(g1), bf16_round(g2)


def keypoints(seed, K, W, H):
    """Integer-valued fp32 pixel keypoints in [3, W-4] x [3, H-4], (K, 2) as (x, y)."""
    g = _gen(seed)
    x = torch.randint(3, W - 3, (K,), generator=g)
    y = torch.randint(3, H - 3, (K,), generator=g)
    return torch.stack([x, y], dim=-1).float()


def points3d(seed, K, noise=0.02):
    g = _gen(seed)
    p1 = torch.rand(K, 3, generator=g)
    p2 = p1 + noise * torch.randn(K, 3, generator=g)
    return p1, p2


def depths(seed, K, lo=0.5, hi=5.0):
    g = _gen(seed)
    return lo + (hi - lo) * torch.rand(K, generator=g)


def head_params(seed, D, hidden=128):
    """Default-initialised ``DepthAwareFeatureFusion.fusion_layer`` parameters under a seed.

    Returned as a dict of fp32 tensors (W1 (hidden, D), b1, gamma, beta (hidden,),
    w2 (1, hidden), b2 (1,)); gamma/beta/biases are perturbed a little so their
    gradients are exercised.
    """
    g = _gen(seed)
    k1 = 1.0 / math.sqrt(D)
    k2 = 1.0 / math.sqrt(hidden)
    return dict(
        W1=(torch.rand(hidden, D, generator=g) * 2 - 1) * k1,
        b1=(torch.rand(hidden, generator=g) * 2 - 1) * k1,
        gamma=1.0 + 0.1 * torch.randn(hidden, generator=g),
        beta=0.1 * torch.randn(hidden, generator=g),
        w2=(torch.rand(1, hidden, generator=g) * 2 - 1) * k2,
        b2=(torch.rand(1, generator=g) * 2 - 1) * k2,
    )


def load_head(head, params):
    """Copy ``head_params`` into a DepthHead / DepthAwareFeatureFusion-shaped module."""
    with torch.no_grad():
        fl = head.fusion_layer
        fl[0].weight.copy_(params['W1'])
        fl[0].bias.copy_(params['b1'])
        fl[1].weight.copy_(params['gamma'])
        fl[1].bias.copy_(params['beta'])
        fl[3].weight.copy_(params['w2'])
        fl[3].bias.copy_(params['b2'])
    return head


@dataclass
class PairInputs:
    """Everything one image pair feeds into the three losses."""
    f1: torch.Tensor
    f2: torch.Tensor
    t12: torch.Tensor
    t21: torch.Tensor
    m1: torch.Tensor
    m2: torch.Tensor
    g1: torch.Tensor          # AP / depth token maps (N, C)
    g2: torch.Tensor
    kp1: torch.Tensor         # (K, 2) pixels
    kp2: torch.Tensor
    p1: torch.Tensor          # (K, 3)
    p2: torch.Tensor
    d1: torch.Tensor          # (K,) depths
    d2: torch.Tensor


def pair_inputs(cfg_id, pair_index, N, C, K, grid, variant, patch=14, mask_mode='bernoulli'):
    """Inputs of pair ``pair_index`` of config ``cfg_id`` (seed = 1000*cfg + pair)."""
    seed = 1000 * cfg_id + pair_index
    ph, pw = grid
    f1, f2 = features(seed * 16 + 0, N, C)
    t12 = teacher_volume(seed * 16 + 1, N, variant)
    t21 = teacher_volume(seed * 16 + 2, N, variant)
    m1 = patch_mask(seed * 16 + 3, N, mode=mask_mode)
    m2 = patch_mask(seed * 16 + 4, N, mode=mask_mode)
    g1, g2 = ap_token_maps(seed * 16 + 5, N, C)
    kp1 = keypoints(seed * 16 + 6, K, pw * patch, ph * patch)
    kp2 = keypoints(seed * 16 + 7, K, pw * patch, ph * patch)
    p1, p2 = points3d(seed * 16 + 8, K)
    d1 = depths(seed * 16 + 9, K)
    d2 = depths(seed * 16 + 10, K)
    return PairInputs(f1, f2, t12, t21, m1, m2, g1, g2, kp1, kp2, p1, p2, d1, d2)


# ---- fast_nn descriptor sets (cfg3) ---------------------------------------------------------

def nn_exact_set(seed, n, dim=24, dup=64):
    """Descriptors with entries k/8, k in [-8, 8]: every dot product is exact in fp32
    whatever the summation order, and ties are frequent.  ``dup`` rows are duplicated."""
    g = _gen(seed)
    X = torch.randint(-8, 9, (n, dim), generator=g).float() / 8.0
    if dup and n > 2 * dup:
        src = torch.randperm(n, generator=g)[:dup]
        dst = torch.randperm(n, generator=g)[:dup]
        X[dst] = X[src]
    return X.contiguous()


def nn_real_set(seed, n, dim=24):
    """Gaussian, L2-normalised fp32 descriptors (MASt3R-like)."""
    g = _gen(seed)
    X = torch.randn(n, dim, generator=g)
    return torch.nn.functional.normalize(X, dim=-1).contiguous()


def nn_desc_maps(seed, H, W, dim=24, noise=0.15):
    """Two correlated dense descriptor maps (H, W, dim) for ``fast_reciprocal_NNs``:
    map 2 is map 1 shifted by a smooth warp plus noise, so reciprocal matches exist."""
    g = _gen(seed)
    coarse = torch.randn(1, dim, H // 8 + 2, W // 8 + 2, generator=g)
    base = torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=True)[0]
    base = base.permute(1, 2, 0) + 0.35 * torch.randn(H, W, dim, generator=g)
    d1 = torch.nn.functional.normalize(base, dim=-1)
    d2 = torch.roll(base, shifts=(3, -5), dims=(0, 1)) + noise * torch.randn(H, W, dim, generator=g)
    d2 = torch.nn.functional.normalize(d2, dim=-1)
    return d1.contiguous(), d2.contiguous()


def point_map(seed, h, w, fx, fy, cx, cy, oversample=1.15, noise=0.004, behind=0.02):
    """A camera-frame point map like the teacher's ``pts3d``: one point per pixel of a slightly denser grid than the
    target image (so that several points can land on one pixel and some pixels stay empty), smooth depth in
    [0.8, 4], a little lateral noise, and a fraction of points pushed behind the camera."""
    g = torch.Generator().manual_seed(seed)
    hh, ww = int(h * oversample), int(w * oversample)
    v, u = torch.meshgrid(torch.linspace(-3, h + 2, hh), torch.linspace(-3, w + 2, ww), indexing='ij')
    z = 2.4 + 1.6 * torch.sin(u / 9.0) * torch.cos(v / 7.0)
    x = (u - cx) / fx * z + noise * torch.randn(hh, ww, generator=g)
    y = (v - cy) / fy * z + noise * torch.randn(hh, ww, generator=g)
    pts = torch.stack([x, y, z], dim=-1).reshape(-1, 3)
    flip = torch.rand(len(pts), generator=g) < behind
    pts[flip, 2] = -pts[flip, 2]
    pts[::97, 2] = 0.0
    return pts.float().contiguous()


def depth_splat_cases():
    """name -> (points (M, 3), K (3, 3), w, h) for ``point_cloud_to_depth``."""
    def intr(fx, fy, cx, cy):
        return torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)

    cases = {}
    cases['scene'] = (point_map(71, 48, 64, 60.0, 58.0, 31.5, 23.5), intr(60.0, 58.0, 31.5, 23.5), 64, 48)
    g = torch.Generator().manual_seed(72)
    sparse = torch.rand(500, 3, generator=g) * torch.tensor([2.0, 1.5, 3.0]) + torch.tensor([-1.0, -0.75, 0.5])
    cases['sparse'] = (sparse.float(), intr(40.0, 40.0, 20.0, 15.0), 40, 30)
    cases['none'] = (torch.cat([sparse[:50, :2], -sparse[:50, 2:]], dim=1).float(), intr(40.0, 40.0, 20.0, 15.0), 40, 30)
    # projections that fall exactly on k + 0.5 (z = 1, power-of-two focal): round-half-even decides the pixel
    k = torch.arange(-2, 18, dtype=torch.float32)
    xs, ys = torch.meshgrid(k / 64.0, k / 64.0, indexing='xy')
    half = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(xs.numel())], dim=1)
    half[:, 2] += 0.0
    cases['halfpix'] = (half.float(), intr(64.0, 64.0, 0.5, 0.5), 16, 16)
    return cases


def vggt_qk(seed, B, heads, n, head_dim=64, skip=5, peak=0.6):
    """bf16 q / k of one VGGT global block for two views of n patch tokens (+ ``skip`` special tokens each): unit-ish
    Gaussian vectors (q_norm / k_norm are LayerNorms), with the patch tokens of view 2 correlated to a permutation of
    view 1 so that the attention rows are peaked like a real teacher's."""
    g = torch.Generator().manual_seed(seed)
    T = 2 * (n + skip)
    q = torch.randn(B, heads, T, head_dim, generator=g)
    k = torch.randn(B, heads, T, head_dim, generator=g)
    perm = torch.randperm(n, generator=g)
    half = T // 2
    k[:, :, half + skip:] += peak * q[:, :, skip:half][:, :, perm]
    k[:, :, skip:half] += peak * q[:, :, half + skip:][:, :, torch.argsort(perm)]
    return q.bfloat16(), k.bfloat16()


def analytic_scene(h, w, view):
    """Smooth per-pixel 3-D points (h, w, 3) and depths (h, w) of ``view`` 0 / 1, computed from pixel coordinates only
    (no random state), for the end-to-end step fixtures: both the golden generator and the tests evaluate this."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    u, v = xs / w, ys / h
    z = 1.5 + 0.8 * torch.sin(3.0 * u + 0.7 * view) * torch.cos(2.0 * v) + 0.05 * view
    pts = torch.stack([(u - 0.5) * z, (v - 0.5) * z, z], dim=-1)
    return pts.contiguous(), z.contiguous()
