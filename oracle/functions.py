"""Oracle restatements of the hot-path helpers in the reference's ``utils/functions.py``.

Test infrastructure only (see ``oracle/__init__.py``).  Plain PyTorch on CPU.
"""

import torch
import torch.nn.functional as F


def sigmoid(x, temp=1.0):
    """Clamped temperature sigmoid.  Follows ``utils/functions.py:24-33``.

    y = 1 / (1 + exp(clamp(-x / temp, -50, 50))).  The clamp makes the gradient
    exactly zero wherever |x / temp| > 50.
    """
    e = (-x / temp).clamp(min=-50, max=50)
    return 1.0 / (1.0 + e.exp())


def patch_grid_affine(h, w, patch_size=14, stride=14):
    """The pixel -> normalised-grid affine map of ``utils/functions.py:56-65``.

    Returns (aw, ah, bw, bh) as python floats, computed with the reference's own
    operation order so the fp64 -> fp32 rounding of the coefficients is the same.
    """
    half = patch_size / 2
    last_h = ((h - patch_size) // stride) * stride + half
    last_w = ((w - patch_size) // stride) * stride + half
    ah = 2 / (last_h - half)
    aw = 2 / (last_w - half)
    bh = 1 - last_h * 2 / (last_h - half)
    bw = 1 - last_w * 2 / (last_w - half)
    return aw, ah, bw, bh


def interpolate_features(descriptors, pts, h, w, normalize=True, patch_size=14, stride=14):
    """Bilinear sampling of an NCHW feature map at pixel keypoints.

    Follows ``utils/functions.py:55-76``: pixel (x, y) is mapped affinely so that
    patch centres land on [-1, 1], then ``grid_sample(align_corners=True,
    padding_mode='border')``; output is (B, C, K), optionally L2-normalised over C.
    """
    aw, ah, bw, bh = patch_grid_affine(h, w, patch_size, stride)
    scale = torch.tensor([[aw, ah]]).to(pts).float()
    shift = torch.tensor([[bw, bh]]).to(pts).float()
    grid = (scale * pts + shift).unsqueeze(-3)           # (B, 1, K, 2)
    out = F.grid_sample(descriptors, grid, align_corners=True, padding_mode='border')
    out = out.squeeze(-2)                                # (B, C, K)
    if normalize:
        out = F.normalize(out, dim=1)
    return out


def extract_kp_depth(depth_map, kp, window_size=3):
    """Mean depth in a replicate-padded window around integer keypoints.

    Follows ``utils/functions.py:348-372``: pad (replicate) -> unfold(window) ->
    mean over the window -> gather at ``y * W + x``.  Returns (B, K).
    """
    if not torch.is_tensor(depth_map):
        depth_map = torch.tensor(depth_map, device=kp.device, dtype=torch.float)
    dm = depth_map[None, None]
    H, W = dm.shape[-2:]
    r = window_size // 2
    padded = F.pad(dm, (r, r, r, r), mode='replicate')
    means = F.unfold(padded, kernel_size=window_size, stride=1).mean(dim=1)   # (1, H*W)
    flat = (kp[..., 1] * W + kp[..., 0]).long()
    return means.gather(dim=1, index=flat)


def get_patch_mask_from_kp_tensor(kp_xy, H, W, patch_size, device=None):
    """Boolean (num_patches,) mask of the patches that contain >= 1 keypoint.

    Follows ``utils/functions.py:375-399``; out-of-image keypoints are dropped,
    patch index = (y // p) * (W // p) + (x // p) on the truncated coordinates.
    """
    if device is None:
        device = kp_xy.device
    ph, pw = H // patch_size, W // patch_size
    mask = torch.zeros(ph * pw, dtype=torch.bool, device=device)
    x, y = kp_xy[:, 0], kp_xy[:, 1]
    inside = (x >= 0) & (x < W) & (y >= 0) & (y < H)
    if inside.sum() == 0:
        return mask
    xi = x[inside].long() // patch_size
    yi = y[inside].long() // patch_size
    mask[yi * pw + xi] = True
    return mask


def get_masked_patch_cost(cost, mask_patch_1, mask_patch_2=None, eps=1e-8,
                          use_softmax=False, temperature=1.0):
    """Zero the masked-out part of a (B, N, N2) cost volume, then normalise rows.

    Follows ``utils/functions.py:402-422``.  Rows with ``mask_patch_1 == False``
    (and columns with ``mask_patch_2 == False`` when given) are set to 0; then
    either ``softmax(./temperature, dtype=float32)`` or ``./clamp_min(rowsum, eps)``.
    """
    B, n1, n2 = cost.shape
    if mask_patch_2 is None:
        keep = mask_patch_1[:, None] & torch.ones(n2, dtype=torch.bool, device=cost.device)[None, :]
    else:
        keep = mask_patch_1[:, None] & mask_patch_2[None, :]
    keep = keep[None].expand(B, n1, n2)
    out = torch.where(keep, cost, torch.zeros((), dtype=cost.dtype, device=cost.device))
    if use_softmax:
        return torch.softmax(out / temperature, dim=-1, dtype=torch.float32)
    return out / out.sum(dim=-1, keepdim=True).clamp_min(eps)


def filter_kp_by_conf(kp, conf_mask):
    """Keep the keypoints whose rounded pixel is True in ``conf_mask``.

    Follows ``utils/functions.py:199-207``.  kp is (1, K, 2); returns
    (kp[:, valid], valid_idx).
    """
    xy = kp[0]
    xi = xy[:, 0].round().long()
    yi = xy[:, 1].round().long()
    idx = conf_mask[yi, xi].nonzero(as_tuple=False).squeeze(1)
    return kp[:, idx, :], idx


def point_cloud_to_depth(points, K, w, h, device=None):
    """Pinhole splat of camera-frame points into a (1, 1, h, w) depth image, averaging the depths that land on the
    same pixel.  Follows ``utils/functions.py:218-260``.

    Only points in front of the camera (z > 0) count; the pixel is round-half-even of (x / z) * fx + cx (three
    separately rounded fp32 operations, no fused multiply-add), points outside the image are dropped, empty
    pixels stay 0.
    """
    pts = points.to(torch.float32)
    K = K.to(torch.float32)
    sums = torch.zeros(h * w, dtype=torch.float32)
    hits = torch.zeros(h * w, dtype=torch.float32)
    front = pts[pts[:, 2] > 0]
    if len(front):
        z = front[:, 2]
        col = torch.round((front[:, 0] / z) * K[0, 0] + K[0, 2]).long()
        row = torch.round((front[:, 1] / z) * K[1, 1] + K[1, 2]).long()
        inside = (col >= 0) & (col < w) & (row >= 0) & (row < h)
        pix = (row * w + col)[inside]
        sums.index_add_(0, pix, z[inside])
        hits.index_add_(0, pix, torch.ones_like(z[inside]))
    depth = torch.where(hits > 0, sums / hits.clamp_min(1), torch.zeros(()))
    return depth.view(1, 1, h, w)
