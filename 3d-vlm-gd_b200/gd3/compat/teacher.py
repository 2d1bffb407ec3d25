"""CUDA replacement for the cost-volume post-processing inside the MASt3R teacher's forward
(``dust3r/dust3r/model.py:346-363``): ``res2['tgt_attn_map'] = teacher_volume(tgt_camap, src_camap,
self.temperature, self.reciprocity)``."""
import torch

from .. import _lib


@torch.no_grad()
def teacher_volume(tgt_camap, src_camap, temperature=3.0, reciprocity=True):
    """Lists of per-layer logits (B, heads, N, N) -> tgt_attn_map (B, N, N), every logit read once."""
    return _lib.teacher_volume(tgt_camap, src_camap, temperature=temperature, reciprocity=reciprocity)
