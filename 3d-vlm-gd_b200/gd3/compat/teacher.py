"""CUDA replacement for the cost-volume post-processing inside the MASt3R teacher's forward
(``dust3r/dust3r/model.py:346-363``): ``res2['tgt_attn_map'] = teacher_volume(tgt_camap, src_camap,
self.temperature, self.reciprocity)``."""
import torch

from .. import _lib


@torch.no_grad()
def teacher_volume(tgt_camap, src_camap, temperature=3.0, reciprocity=True):
    """Lists of per-layer logits (B, heads, N, N) -> tgt_attn_map (B, N, N), every logit read once."""
    return _lib.teacher_volume(tgt_camap, src_camap, temperature=temperature, reciprocity=reciprocity)


@torch.no_grad()
def vggt_cost_volumes(attn_list):
    """VGGT teacher: ``attn_mean = torch.mean(torch.stack(attn_list), dim=0)`` (``vggt/models/aggregator.py:273``),
    ``cost_1, cost_2 = attn.chunk(2, dim=0)`` and ``cost_k.mean(dim=1)`` (``src/finetune_timm_vggt.py:390-392``) in one
    pass over the per-block maps.  attn_list: list of (2 B, heads, n, n) -> (cost_1, cost_2), each (B, n, n)."""
    out = _lib.teacher_volume(attn_list, None, plain_mean=True)
    return out.chunk(2, dim=0)
