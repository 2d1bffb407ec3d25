"""Teacher-side keypoints on the device: the body of ``FinetuneMASt3RTIMM.filter_and_match_keypoints``
(``src/finetune_timm_mast3r.py:392-469``) without its host round trips.

The reference runs ``fast_reciprocal_NNs`` (one device -> host copy per block query, numpy bookkeeping), filters the
matches on the host and uploads the survivors.  Here the seed ping-pong is ``gd3_fast_reciprocal_nn`` (seed state
resident on the GPU), and the unique / sort of ``merge_corres``, the border filter, the confidence-quantile masks and
``filter_kp_by_conf`` are device-side torch expressions on its output; the keypoints never leave the GPU.
"""
import numpy as np
import torch

from .. import _lib


@torch.no_grad()
def filter_and_match_keypoints(mast3r_features, min_conf_thr, subsample=16, border=3, device='cuda'):
    """-> (kp_1, kp_2, w, h): matched pixel keypoints (1, n, 2) fp32 (x, y) of both views on ``device``, or four
    ``None`` when nothing survives.  ``mast3r_features``: ``desc_1`` / ``desc_2`` (H, W, D) descriptor maps and
    ``conf_1`` / ``conf_2`` (H, W) confidences, as ``extract_mast3r_features`` returns them."""
    desc1, desc2 = mast3r_features['desc_1'], mast3r_features['desc_2']
    conf1, conf2 = mast3r_features['conf_1'], mast3r_features['conf_2']
    H1, W1, D = desc1.shape
    H2, W2, _ = desc2.shape
    _lib.require_cuda()
    flat1 = desc1.to(device).reshape(-1, D).contiguous().float()
    flat2 = desc2.to(device).reshape(-1, D).contiguous().float()
    ys, xs = np.mgrid[subsample // 2:H1:subsample, subsample // 2:W1:subsample].reshape(2, -1)
    seeds = torch.from_numpy(np.int32(np.unique(xs + W1 * ys))).to(flat1.device)
    if seeds.numel() == 0:
        return None, None, None, None
    i1, i2, converged = _lib.fast_reciprocal_nn(flat1, flat2, seeds, max_iter=10, dist='dot')
    # merge_corres (mast3r/fast_nn.py:87-106): unique (idx1, idx2) pairs ordered by idx1 then idx2
    key = torch.unique((i1[converged].long() << 32) | i2[converged].long())
    i1, i2 = key >> 32, key & 0xFFFFFFFF
    kp1 = torch.stack([i1 % W1, i1 // W1], dim=-1)
    kp2 = torch.stack([i2 % W2, i2 // W2], dim=-1)

    def inside(kp, w, h):
        return (kp[:, 0] >= border) & (kp[:, 0] < w - border) & (kp[:, 1] >= border) & (kp[:, 1] < h - border)

    keep = inside(kp1, W1, H1) & inside(kp2, W2, H2)
    kp1, kp2 = kp1[keep], kp2[keep]

    def confident(kp, conf, h, w):
        # pixels at or above the min_conf_thr-th percentile of the view's confidences (:444-452)
        c = conf.to(flat1.device).reshape(-1)
        thr = c.sort()[0][int(c.shape[0] * float(min_conf_thr) * 0.01)]
        return (c.reshape(h, w) >= thr)[kp[:, 1], kp[:, 0]]

    # a match stays when either of its end points is confident (union of the two index sets, :454-460)
    keep = confident(kp1, conf1, H1, W1) | confident(kp2, conf2, H2, W2)
    kp1, kp2 = kp1[keep].float()[None], kp2[keep].float()[None]
    if kp1.shape[1] == 0:
        return None, None, None, None
    return kp1, kp2, int(W1), int(H1)
