"""Oracle restatements of the loss *bodies* inside the reference's LightningModules.

The reference computes the three distillation losses inline in
``calculate_cost_loss`` / ``calculate_matching_loss`` / ``calculate_depth_loss``
(``src/finetune_timm_{mast3r,vggt,me}.py``).  Those modules cannot be imported
here (timm / pytorch_lightning / hydra are absent), so each body is restated
around plain tensors, calling the oracle versions of the helpers the reference
calls.  One image pair per call, exactly like the reference (B = 1).

Test infrastructure only (see ``oracle/__init__.py``).
"""

import torch
import torch.nn.functional as F

from .functions import get_masked_patch_cost, interpolate_features, sigmoid
from .losses import kl_divergence_map, pairwise_logistic_ranking_loss


def cost_volume_kl(f1, f2, teacher12, teacher21, mask1, mask2, variant='mast3r', eps=1e-8):
    """Dense cost-volume KL for ONE pair.

    f1, f2: (N, C) student patch features (un-normalised); teacher12/21: (N, N)
    teacher volumes (rows = patches of view 1 / view 2); mask1/2: (N,) bool.

    variant 'mast3r' follows ``src/finetune_timm_mast3r.py:521-540``: masked rows
    of the student logits are zeroed *before* a float32 softmax.
    variant 'vggt' follows ``src/finetune_timm_vggt.py:510-533``: softmax on all
    rows first, then masked rows zeroed and rows re-normalised by their sum.
    """
    a = F.normalize(f1[None], p=2, dim=-1)
    b = F.normalize(f2[None], p=2, dim=-1)
    z12 = torch.bmm(a, b.transpose(-1, -2))
    z21 = torch.bmm(b, a.transpose(-1, -2))
    t12 = get_masked_patch_cost(teacher12[None], mask1, None, eps=eps)
    t21 = get_masked_patch_cost(teacher21[None], mask2, None, eps=eps)
    if variant == 'mast3r':
        s12 = get_masked_patch_cost(z12, mask1, None, use_softmax=True)
        s21 = get_masked_patch_cost(z21, mask2, None, use_softmax=True)
    elif variant == 'vggt':
        s12 = get_masked_patch_cost(torch.softmax(z12, dim=-1), mask1, None)
        s21 = get_masked_patch_cost(torch.softmax(z21, dim=-1), mask2, None)
    else:
        raise ValueError(f'unknown {variant=}')
    return (kl_divergence_map(t12, s12, eps) + kl_divergence_map(t21, s21, eps)) / 2


def smooth_ap(d1, d2, pts3d_1, pts3d_2, variant='mast3r', temp=0.01, thr_neg=0.1, thr_pos=5e-3):
    """Smooth-AP sparse-correspondence loss for ONE pair.

    d1, d2: (K, C) L2-normalised descriptors; pts3d_*: (K, 3).

    'mast3r' follows ``src/finetune_timm_mast3r.py:557-589`` (positives on the
    diagonal, r1 = 1 + sig(pos - 1)); 'vggt' follows
    ``src/finetune_timm_vggt.py:543-574`` (r1 = 1 + sig(1 - pos)); 'me' follows
    ``src/finetune_timm_me.py:196-217`` (positives = every (s, t) closer than
    ``thr_pos`` in 3-D, negatives = farther than ``thr_neg``, no diagonal rule).
    """
    dist = torch.cdist(pts3d_1[None], pts3d_2[None])[0]
    sim = d1 @ d2.transpose(-1, -2)
    K = d1.shape[0]
    if variant in ('mast3r', 'vggt'):
        rows = torch.arange(K, device=d1.device)
        cols = rows
        neg = (dist > thr_neg) & ~torch.eye(K, dtype=torch.bool, device=d1.device)
    elif variant == 'me':
        rows, cols = torch.nonzero(dist < thr_pos, as_tuple=True)
        neg = dist > thr_neg
    else:
        raise ValueError(f'unknown {variant=}')
    pos = sim[rows, cols]
    sim_r = sim[rows]
    neg_r = neg[rows].float()
    if variant == 'vggt':
        r1 = sigmoid(1.0 - pos, temp) + 1
    else:
        r1 = sigmoid(pos - 1.0, temp) + 1
    ap1 = r1 / (r1 + (sigmoid(sim_r - 1.0, temp) * neg_r).sum(dim=-1))
    r2 = sigmoid(1.0 - pos, temp) + 1
    ap2 = r2 / (r2 + (sigmoid(sim_r - pos[:, None], temp) * neg_r).sum(dim=-1))
    return torch.mean(1.0 - (ap1 + ap2) / 2)


def smooth_ap_me_batch(d1, d2, pts3d_1, pts3d_2, temp=0.01, thr_neg=0.1, thr_pos=5e-3):
    """The ME baseline's loss on a batch of B pairs, ``src/finetune_timm_me.py:199-217`` line for line: the positives
    of ALL pairs are gathered by one ``nonzero`` and ONE mean is taken over them (so pairs with many positives weigh
    more, and a pair without positives simply does not count).  d1, d2: (B, K, C); pts3d_*: (B, K, 3)."""
    dist = torch.cdist(pts3d_1, pts3d_2)
    sim = torch.bmm(d1, d2.transpose(-1, -2))
    pos_idxs = torch.nonzero(dist < thr_pos, as_tuple=False)
    pos_sim = sim[pos_idxs[:, 0], pos_idxs[:, 1], pos_idxs[:, 2]]
    rpos = sigmoid(pos_sim - 1.0, temp) + 1
    neg_mask = dist[pos_idxs[:, 0], pos_idxs[:, 1]] > thr_neg
    sim_rows = sim[pos_idxs[:, 0], pos_idxs[:, 1]]
    ap1 = rpos / (rpos + torch.sum(sigmoid(sim_rows - 1.0, temp) * neg_mask.float(), -1))
    rpos = sigmoid(1.0 - pos_sim, temp) + 1
    ap2 = rpos / (rpos + torch.sum(sigmoid(sim_rows - pos_sim[:, None], temp) * neg_mask.float(), -1))
    return torch.mean(1.0 - (ap1 + ap2) / 2)


def sample_tokens(tokens, ph, pw, kp, patch_size=14, stride=14, normalize=False):
    """Sample token-major features (P, N, C) at pixel keypoints (P, K, 2) -> (P, K, C).

    This is the reference's call pattern around ``interpolate_features``:
    ``reshape(B, ph, pw, C).permute(0, 3, 1, 2)`` -> ``interpolate_features(...,
    h=ph*patch, w=pw*patch, normalize=False).permute(0, 2, 1)`` -> optional
    ``F.normalize(dim=-1)`` (``src/finetune_timm_mast3r.py:271-274,307-313``).
    """
    P, N, C = tokens.shape
    fmap = tokens.reshape(P, ph, pw, C).permute(0, 3, 1, 2).contiguous()
    out = interpolate_features(fmap, kp, h=ph * patch_size, w=pw * patch_size,
                               patch_size=patch_size, stride=stride, normalize=False).permute(0, 2, 1)
    if normalize:
        out = F.normalize(out, p=2, dim=-1)
    return out


def depth_losses(head, kp_feat_1, kp_feat_2, kp_depth_1, kp_depth_2, depth_threshold=0.05):
    """Cross-view L1 + intra-view ranking for ONE pair.

    kp_feat_*: (1, K, D); kp_depth_*: (1, K).  Follows
    ``src/finetune_timm_mast3r.py:489-499`` (identical in vggt ``:473-483``).
    Returns (depth_loss, intra_depth_loss).
    """
    pred = head(kp_feat_1 - kp_feat_2)
    depth_loss = F.l1_loss(pred, torch.tanh(kp_depth_1 - kp_depth_2).detach())
    r1 = pairwise_logistic_ranking_loss(head, kp_feat_1, kp_depth_1, depth_threshold=depth_threshold)
    r2 = pairwise_logistic_ranking_loss(head, kp_feat_2, kp_depth_2, depth_threshold=depth_threshold)
    return depth_loss, (r1 + r2) / 2
