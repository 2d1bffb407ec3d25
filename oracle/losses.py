"""Oracle restatements of ``utils/losses.py`` and of the depth head it is called with.

Test infrastructure only (see ``oracle/__init__.py``).  Plain PyTorch on CPU.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F


class DepthHead(nn.Module):
    """Parameter-compatible restatement of ``DepthAwareFeatureFusion``.

    Follows ``utils/model.py:88-127``.  Only the branch the losses use
    (``depths=None``: ``fusion_layer`` then optional tanh) is implemented; the
    ``depth_attention`` branch is kept as parameters so state dicts line up.
    """

    def __init__(self, input_dim, hidden_dim=128, use_tanh=True):
        super().__init__()
        self.use_tanh = use_tanh
        self.depth_attention = nn.Sequential(
            nn.Linear(1, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, input_dim), nn.Sigmoid())
        self.fusion_layer = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.GELU(),
            nn.Linear(hidden_dim, 1))

    def forward(self, features, depths=None):
        if depths is not None:
            features = features * self.depth_attention(depths.unsqueeze(-1))
        out = self.fusion_layer(features)
        if self.use_tanh:
            out = torch.tanh(out)
        return out.squeeze(-1)


def kl_divergence_map(teacher, student, eps=1e-8):
    """KL(teacher || student) summed over columns, averaged over all B*N rows.

    Follows ``utils/losses.py:5-15``: both operands are clamped at ``eps`` first.
    """
    t = teacher.clamp_min(eps)
    s = student.clamp_min(eps)
    return (t * torch.log(t / s)).sum(dim=-1).mean()


def pairwise_logistic_ranking_loss(model, feats, depths, depth_threshold=0.0):
    """Pairwise logistic ranking loss on the head's score of feature differences.

    Follows ``utils/losses.py:18-41``: for every ordered pair (i, j):
    s_ij = model(f_j - f_i), alpha_ij = sign(d_j - d_i), loss_ij =
    log(1 + exp(-alpha_ij * s_ij)); mean over pairs with |d_j - d_i| > threshold;
    constant 0 (no grad) when no pair is valid.
    """
    B, K, D = feats.shape
    diff = feats[:, None, :, :] - feats[:, :, None, :]       # [b, i, j] = f_j - f_i
    dd = depths[:, None, :] - depths[:, :, None]             # d_j - d_i
    score = model(diff.reshape(B, K * K, D)).reshape(B, K, K)
    per_pair = torch.log(1.0 + torch.exp(-torch.sign(dd) * score))
    valid = dd.abs() > depth_threshold
    picked = per_pair[valid]
    if picked.numel() == 0:
        return torch.tensor(0.0, device=feats.device)
    return picked.mean()


def intra_depth_loss(model, feats, depths, base_margin=0.05, depth_thresh=0.05):
    """Hinge sibling of the ranking loss.  Follows ``utils/losses.py:44-69``.

    s_ij = model(f_i - f_j); target = sign(tanh(d_i - d_j)); loss_ij =
    relu(margin - target * s_ij) over pairs with |tanh(d_i - d_j)| > thresh.
    """
    B, K, D = feats.shape
    diff = feats[:, :, None, :] - feats[:, None, :, :]       # f_i - f_j
    score = model(diff.reshape(B, K * K, D)).reshape(B, K, K)
    gt = torch.tanh(depths[:, :, None] - depths[:, None, :]).detach()
    per_pair = F.relu(base_margin - torch.sign(gt) * score)
    valid = gt.abs() > depth_thresh
    if valid.sum() > 0:
        return per_pair[valid].mean()
    return torch.tensor(0.0, device=feats.device)


def infonce(desc1, desc2, valid_matches=None, temperature=0.07, eps=1e-8, mode='all'):
    """Upstream MASt3R InfoNCE (softmax-CE correspondence loss), mean-reduced.

    Follows ``mast3r/losses.py:237-272`` (+ ``get_similarities :202-209`` and the
    'mean' reduction of ``MatchingCriterion.forward :217-231``).  No ``src/``
    script calls it; it is here because BASELINE.json words the correspondence
    loss as "softmax-CE" (SURVEY.md section 8, row a3b).
    """
    B, K, _ = desc1.shape
    if valid_matches is None:
        valid_matches = torch.ones(B, K, dtype=torch.bool)
    sim = desc1.float() @ desc2.float().transpose(-2, -1) / temperature
    sim = torch.where(sim.isnan(), torch.full_like(sim, float('-inf')), sim).exp()
    pos = sim.diagonal(dim1=-2, dim2=-1)
    row = sim.sum(dim=-1)
    col = sim.sum(dim=-2)
    if mode == 'all':
        loss = -torch.log((pos / row.sum(dim=-1, keepdim=True)).clip(eps))
    elif mode == 'proper':
        loss = -(torch.log((pos / col).clip(eps)) + torch.log((pos / row).clip(eps)))
    elif mode == 'dual':
        loss = -torch.log((pos ** 2 / row / col).clip(eps))
    else:
        raise ValueError(f'bad {mode=}')
    picked = loss[valid_matches]
    return picked.mean() if picked.numel() > 0 else picked.new_zeros(())
