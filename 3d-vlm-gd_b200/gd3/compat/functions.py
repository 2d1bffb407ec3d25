"""CUDA drop-in for the hot-path subset of the reference's ``utils/functions.py``.

Same names, argument order, defaults and return shapes.  Sampling, the keypoint patch mask and the keypoint depth run
hand-written kernels through lib3dgd.so, and so do ``point_cloud_to_depth`` and ``get_masked_patch_cost`` (forward and
backward); ``sigmoid`` / ``filter_kp_by_conf`` are plain device-side torch expressions (SURVEY.md section 8, row a7:
"negligible; ... or leave in torch").
"""
import torch

from .. import _lib, ops
from .._lib import require_cuda


def sigmoid(tensor, temp=1.0):
    """Clamped temperature sigmoid, ``utils/functions.py:24-33``.

    Stand-alone elementwise helper; inside the fused Smooth-AP kernel (``ops.smooth_ap``) the same
    expression is evaluated in registers.
    """
    exponent = torch.clamp(-tensor / temp, min=-50, max=50)
    return 1.0 / (1.0 + torch.exp(exponent))


def interpolate_features(descriptors, pts, h, w, normalize=True, patch_size=14, stride=14):
    """``utils/functions.py:55-76``: (B, C, h', w') NCHW features sampled at pixel pts (B, K, 2) -> (B, C, K)."""
    require_cuda(descriptors, pts)
    return ops.interpolate_nchw(descriptors, pts, int(h), int(w), int(patch_size), int(stride), bool(normalize))


def get_patch_mask_from_kp_tensor(kp_xy, H, W, patch_size, device=None):
    """``utils/functions.py:375-399``: (K, 2) pixel keypoints -> bool (num_patches,).  One kernel launch, no host sync."""
    require_cuda(kp_xy)
    mask, _ = _lib.kp_prepare(kp_xy[None], int(H), int(W), patch_size=int(patch_size))
    mask = mask[0]
    return mask if device is None else mask.to(device)


def extract_kp_depth(depth_map, kp, window_size=3):
    """``utils/functions.py:348-372``: mean depth in a replicate-padded window at flat index ``(y * W + x).long()``
    (fp32 arithmetic, so fractional keypoints land where the reference's gather lands) -> (B, K).

    Difference: the reference's ``gather`` raises for an index outside the map; raising needs a host sync, so such a
    keypoint gets NaN here (callers filter keypoints to the image first, src/finetune_timm_mast3r.py:421-426)."""
    if not torch.is_tensor(depth_map):
        depth_map = torch.tensor(depth_map, device=kp.device, dtype=torch.float)
    require_cuda(depth_map, kp)
    H, W = depth_map.shape[-2:]
    _, kd = _lib.kp_prepare(kp, int(H), int(W), depth=depth_map, window=int(window_size))
    return kd.to(depth_map.dtype)


class _MaskedPatchCost(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost, mask1, mask2, eps, use_softmax, temperature):
        out, row_sum, m1, m2 = _lib.masked_patch_cost(cost, mask1, mask2, use_softmax, eps, temperature,
                                                      want_row_sum=ctx.needs_input_grad[0])
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(out, row_sum, m1, *(() if m2 is None else (m2,)))
            ctx.cfg = (bool(use_softmax), float(eps), float(temperature), cost.dtype)
        return out if (use_softmax or cost.dtype == torch.float32) else out.to(cost.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        out, row_sum, m1, *rest = ctx.saved_tensors
        use_softmax, eps, temperature, dtype = ctx.cfg
        g = _lib.masked_patch_cost_backward(grad_out, out, row_sum, m1, rest[0] if rest else None, use_softmax, eps,
                                            temperature)
        return g.to(dtype), None, None, None, None, None


def get_masked_patch_cost(cost, mask_patch_1, mask_patch_2=None, eps=1e-8, use_softmax=False, temperature=1.0):
    """``utils/functions.py:402-422`` on an already materialised (B, N, N2) volume: rows of ``mask_patch_1 == False``
    (and columns of ``mask_patch_2 == False``) are overwritten with 0, then every row is divided by its clamped sum or
    goes through ``softmax(x / temperature)`` in fp32.

    One CUDA kernel forward (``gd3_masked_patch_cost``) and one backward instead of clone + masked fill + softmax /
    sum / clamp / divide.  Kept for callers that hold a volume; the fused path (``ops.cost_volume_kl``) never builds one.
    """
    require_cuda(cost, mask_patch_1, mask_patch_2)
    if cost.dim() != 3:
        raise ValueError(f'get_masked_patch_cost: expected a (B, hw, hw2) volume, got {tuple(cost.shape)}')
    return _MaskedPatchCost.apply(cost, mask_patch_1, mask_patch_2, float(eps), bool(use_softmax), float(temperature))


def filter_kp_by_conf(kp, conf_mask):
    """``utils/functions.py:199-207``."""
    xy = kp[0]
    valid = conf_mask[xy[:, 1].round().long(), xy[:, 0].round().long()]
    idx = valid.nonzero(as_tuple=False).squeeze(1)
    return kp[:, idx, :], idx


def point_cloud_to_depth(points, K, w, h, device):
    """``utils/functions.py:218-260``: (N, 3) camera-frame points + (3, 3) intrinsics -> (1, 1, h, w) fp32 depth image
    (mean z per pixel, 0 where no point lands).  A (B, N, 3) batch gives (B, 1, h, w)."""
    require_cuda(points)
    batched = points.dim() == 3
    pts = points if batched else points[None]
    depth = _lib.point_cloud_to_depth(pts, torch.as_tensor(K, dtype=torch.float32, device=points.device), int(w), int(h))
    return depth[:, None].to(device)
