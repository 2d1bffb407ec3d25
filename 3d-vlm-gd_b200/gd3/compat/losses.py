"""CUDA drop-in for the reference's ``utils/losses.py`` (same names, arguments and defaults)."""
import torch

from .. import ops
from .._lib import require_cuda


def kl_divergence_map(mast3r_cost, feat_cost_sim, eps=1e-8):
    """``utils/losses.py:5-15`` on already materialised volumes.

    Kept for callers that hold (B, N, N) volumes.  The training path should call
    ``gd3.ops.cost_volume_kl`` instead, which fuses normalisation, the N x N contraction, masking,
    softmax and this reduction and never builds either volume.
    """
    t = mast3r_cost.clamp_min(eps)
    s = feat_cost_sim.clamp_min(eps)
    return (t * torch.log(t / s)).sum(dim=-1).mean()


def pairwise_logistic_ranking_loss(model, pred_scores, gt_depths, depth_threshold=0.0):
    """``utils/losses.py:18-41``: pred_scores (B, N, D) keypoint features, gt_depths (B, N).

    ``model`` is the depth-difference head (``DepthAwareFeatureFusion``); its ``fusion_layer``
    parameters receive gradients.  One mean over the valid pairs of all B sets; 0 when none is valid.
    """
    require_cuda(pred_scores, gt_depths)
    total, _, _ = ops.depth_head_loss(model, pred_scores, gt_depths, mode='logistic',
                                      depth_threshold=depth_threshold, joint_mean=True)
    return total


def intra_depth_loss(model, kp_feat, kp_depth, base_margin=0.05, depth_thresh=0.05):
    """``utils/losses.py:44-69`` (hinge sibling of the ranking loss)."""
    require_cuda(kp_feat, kp_depth)
    total, _, _ = ops.depth_head_loss(model, kp_feat, kp_depth, mode='hinge', depth_threshold=depth_thresh,
                                      margin=base_margin, joint_mean=True)
    return total
