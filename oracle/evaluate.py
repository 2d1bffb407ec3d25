"""Oracle restatement of the keypoint-transfer block of the reference's evaluation.

Test infrastructure only (see ``oracle/__init__.py``).  Plain PyTorch on CPU.

Pinned: ``oracle/gen_live_bodies.py --eval`` runs the reference's ``semantic_transfer`` unmodified on CPU (dataset
loader, ViT and ``Tensor.cuda`` replaced by stand-ins) and records the ``nn_idx`` it computes
(``tests/golden/eval_argmax.npz``).
"""
import torch
import torch.nn.functional as F


def semantic_argmax(img1_kp_desc, img2_desc, img_size, patch_size=14, stride=14):
    """Follows ``src/evaluate_timm.py:532-547`` line by line.

    img1_kp_desc: (1, C, K) keypoint descriptors of image 1 (``interpolate_features`` output);
    img2_desc: (1, C, ph, pw) patch descriptors of image 2.  The reference pads with
    ``torchvision.transforms.functional.pad(..., padding_mode='edge')`` given as (left, top,
    right, bottom) = (p/2, p/2, S - H - p/2, S - W - p/2); ``F.pad(mode='replicate')`` takes
    (left, right, top, bottom).  Returns (nn_idx (K,), sim (K, S*S)).
    """
    ds_size = ((img_size - patch_size) // stride) * stride + 1
    d2 = F.interpolate(img2_desc, size=(ds_size, ds_size), mode='bilinear', align_corners=True)
    left = top = patch_size // 2
    right = img_size - d2.shape[2] - (patch_size // 2)
    bottom = img_size - d2.shape[3] - (patch_size // 2)
    d2 = F.pad(d2, (left, right, top, bottom), mode='replicate')
    sim = torch.einsum('nfk,nif->nki', img1_kp_desc, d2.permute(0, 2, 3, 1).reshape(1, img_size * img_size, -1))[0]
    nn_idx = torch.argmax(sim, dim=1)
    return nn_idx, sim
