"""CUDA drop-in for the reference's ``utils/losses.py`` (same names, arguments and defaults)."""
import torch

from .. import _lib, ops
from .._lib import require_cuda


class _KLDivergenceMap(torch.autograd.Function):
    """One streaming kernel computes the loss and both gradients; backward scales the stashed gradients."""

    @staticmethod
    def forward(ctx, teacher, student, eps):
        need_t, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        loss, gs, gt = _lib.kl_divergence_map(teacher, student, eps, need_s, need_t)
        ctx.save_for_backward(*(g for g in (gt, gs) if g is not None))
        ctx.have = (gt is not None, gs is not None)
        ctx.dtypes = (teacher.dtype, student.dtype)
        ctx.shapes = (teacher.shape, student.shape)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        saved = list(ctx.saved_tensors)
        gt = saved.pop(0) if ctx.have[0] else None
        gs = saved.pop(0) if ctx.have[1] else None
        out = []
        for g, dt, shp in ((gt, ctx.dtypes[0], ctx.shapes[0]), (gs, ctx.dtypes[1], ctx.shapes[1])):
            out.append(None if g is None else (g * grad_out).to(dt).view(shp))
        return out[0], out[1], None


def kl_divergence_map(mast3r_cost, feat_cost_sim, eps=1e-8):
    """``utils/losses.py:5-15`` on already materialised volumes: mean over rows of sum_j t~ log(t~ / s~) with both
    volumes clamped at ``eps``.

    One fused CUDA pass (``gd3_kl_divergence_map``: loss and the gradients of both arguments; fp32) instead of six
    elementwise torch kernels and their saved N x N tensors.  The training path should still call
    ``gd3.ops.cost_volume_kl``, which fuses normalisation, the N x N contraction, masking, softmax and this reduction
    and never builds either volume.
    """
    require_cuda(mast3r_cost, feat_cost_sim)
    return _KLDivergenceMap.apply(mast3r_cost, feat_cost_sim, float(eps))


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
