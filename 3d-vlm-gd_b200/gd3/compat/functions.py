"""CUDA drop-in for the hot-path subset of the reference's ``utils/functions.py``.

Same names, argument order, defaults and return shapes.  The heavy functions run hand-written
kernels through lib3dgd.so; tiny index helpers are expressed with device-side torch indexing
("negligible; ... or leave in torch").
"""
import torch

from .. import ops
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
    """``utils/functions.py:375-399``: (K, 2) pixel keypoints -> bool (num_patches,)."""
    if device is None:
        device = kp_xy.device
    ph, pw = H // patch_size, W // patch_size
    mask = torch.zeros(ph * pw, dtype=torch.bool, device=device)
    x, y = kp_xy[:, 0], kp_xy[:, 1]
    inside = (x >= 0) & (x < W) & (y >= 0) & (y < H)
    idx = (y.long() // patch_size) * pw + (x.long() // patch_size)
    idx = torch.where(inside, idx, torch.zeros_like(idx))
    # scatter without a host sync: out-of-image keypoints contribute False
    mask.index_put_((idx.to(device),), inside.to(device), accumulate=True)
    return mask


def extract_kp_depth(depth_map, kp, window_size=3):
    """``utils/functions.py:348-372``: mean depth in a replicate-padded window at integer keypoints -> (B, K)."""
    if not torch.is_tensor(depth_map):
        depth_map = torch.tensor(depth_map, device=kp.device, dtype=torch.float)
    H, W = depth_map.shape[-2:]
    half = window_size // 2
    x = kp[..., 0].long()
    y = kp[..., 1].long()
    acc = torch.zeros(kp.shape[:2], dtype=depth_map.dtype, device=kp.device)
    for dy in range(-half, half + 1):
        for dx in range(-half, half + 1):
            acc = acc + depth_map[(y + dy).clamp(0, H - 1), (x + dx).clamp(0, W - 1)]
    return acc / float(window_size * window_size)


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
