"""Oracle restatement of the teacher-side keypoint selection of the MASt3R fine-tuning module.

Test infrastructure only (see ``oracle/__init__.py``).  Plain PyTorch / numpy on CPU.  Pinned against the live
``FinetuneMASt3RTIMM.filter_and_match_keypoints`` (``tests/golden/live_bodies.npz``, cases ``kpmatch*``).
"""
import torch

from . import fast_nn
from .functions import filter_kp_by_conf


def filter_and_match_keypoints(desc_1, desc_2, conf_1, conf_2, min_conf_thr, subsample=16, border=3):
    """Follows ``src/finetune_timm_mast3r.py:392-469``.

    Reciprocal matches from a ``subsample`` seed grid (``:416-419``), matches with an end point closer than ``border``
    pixels to the image edge dropped (``:422-433``), then a match is kept when either end point lies on a pixel whose
    confidence reaches the ``min_conf_thr``-th percentile of its view (``:444-460``).
    Returns (kp_1, kp_2) as (1, n, 2) fp32 (x, y), or (None, None).
    """
    H1, W1 = desc_1.shape[:2]
    H2, W2 = desc_2.shape[:2]
    xy1, xy2 = fast_nn.fast_reciprocal_NNs(desc_1, desc_2, subsample_or_initxy1=subsample, device='cpu', dist='dot',
                                           block_size=2 ** 13)
    ok = ((xy1[:, 0] >= border) & (xy1[:, 0] < W1 - border) & (xy1[:, 1] >= border) & (xy1[:, 1] < H1 - border)
          & (xy2[:, 0] >= border) & (xy2[:, 0] < W2 - border) & (xy2[:, 1] >= border) & (xy2[:, 1] < H2 - border))
    kp_1 = torch.tensor(xy1[ok]).float()[None]
    kp_2 = torch.tensor(xy2[ok]).float()[None]

    def conf_mask(conf, h, w):
        ordered = conf.reshape(-1).sort()[0]
        return conf.reshape(h, w) >= ordered[int(ordered.shape[0] * float(min_conf_thr) * 0.01)]

    _, keep_1 = filter_kp_by_conf(kp_1, conf_mask(conf_1, H1, W1))
    _, keep_2 = filter_kp_by_conf(kp_2, conf_mask(conf_2, H2, W2))
    keep = torch.unique(torch.cat([keep_1, keep_2]))
    kp_1, kp_2 = kp_1[:, keep], kp_2[:, keep]
    if kp_1.shape[1] == 0:
        return None, None
    return kp_1, kp_2
