"""CUDA drop-in for the hot-path subset of the reference's ``utils/functions.py``.

Same names, argument order, defaults and return shapes.  Sampling, the keypoint patch mask and the keypoint depth run
hand-written kernels through lib3dgd.so, and so does ``point_cloud_to_depth``; ``sigmoid`` / ``get_masked_patch_cost`` / ``filter_kp_by_conf`` are kept for
callers that still hold materialised tensors and are plain device-side torch expressions (SURVEY.md section 8, row
a7: "negligible; ... or leave in torch").
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


def get_masked_patch_cost(cost, mask_patch_1, mask_patch_2=None, eps=1e-8, use_softmax=False, temperature=1.0):
    """``utils/functions.py:402-422`` on an already materialised (B, N, N2) volume.

    Kept for callers that hold a volume; the fused path (``ops.cost_volume_kl``) never builds one.
    """
    B, n1, n2 = cost.shape
    keep = mask_patch_1[:, None] if mask_patch_2 is None else (mask_patch_1[:, None] & mask_patch_2[None, :])
    out = torch.where(keep[None].expand(B, n1, n2), cost, cost.new_zeros(()))
    if use_softmax:
        return torch.softmax(out / temperature, dim=-1, dtype=torch.float32)
    return out / out.sum(dim=-1, keepdim=True).clamp_min(eps)


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
