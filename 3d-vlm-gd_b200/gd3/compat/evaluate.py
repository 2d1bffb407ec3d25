"""CUDA replacement for the keypoint-transfer block of ``src/evaluate_timm.py:532-547``.

The reference code is inline in its evaluation loop (not a function); the helper below takes the
same tensors that block works on and returns the same ``nn_idx`` / ``kps_1_to_2``, without
building the upsampled (1, C, img, img) descriptor map.
"""
import torch

from .. import _lib


@torch.no_grad()
def semantic_argmax(img1_kp_desc, img2_desc, img_size, patch_size=14, stride=14):
    """img1_kp_desc: (1, C, K) from ``interpolate_features(..., normalize=True)``; img2_desc: (1, C, ph, pw).

    Returns (nn_idx (K,), kps_1_to_2 (K, 2) as (x, y)) exactly like lines :542-547."""
    nn_idx = _lib.semantic_argmax(img1_kp_desc, img2_desc, img_size, patch_size, stride)
    nn_x = nn_idx % img_size
    nn_y = nn_idx // img_size
    return nn_idx, torch.stack([nn_x, nn_y]).permute(1, 0)
