"""CUDA replacement for the cost-volume post-processing inside the MASt3R teacher's forward
(``dust3r/dust3r/model.py:346-363``): ``res2['tgt_attn_map'] = teacher_volume(tgt_camap, src_camap,
self.temperature, self.reciprocity)``."""
import torch

from .. import _lib


@torch.no_grad()
def teacher_volume(tgt_camap, src_camap, temperature=3.0, reciprocity=True, packed=False, eps=1e-8):
    """Lists of per-layer logits (B, heads, N, N) -> tgt_attn_map (B, N, N), every logit read once.

    packed=True hands the volume over in the form ``gd3.ops.cost_volume_kl`` reads fastest: ``(fp16 volume * 1024,
    row statistics (B, 3, N))`` (``gd3.ops.pack_teacher``: half the bytes, no statistics pass in the loss)."""
    out = _lib.teacher_volume(tgt_camap, src_camap, temperature=temperature, reciprocity=reciprocity)
    if packed:
        from .. import ops
        return ops.pack_teacher(out, eps=eps)
    return out


@torch.no_grad()
def vggt_cost_volumes(attn_list):
    """VGGT teacher: ``attn_mean = torch.mean(torch.stack(attn_list), dim=0)`` (``vggt/models/aggregator.py:273``),
    ``cost_1, cost_2 = attn.chunk(2, dim=0)`` and ``cost_k.mean(dim=1)`` (``src/finetune_timm_vggt.py:390-392``) in one
    pass over the per-block maps.  attn_list: list of (2 B, heads, n, n) -> (cost_1, cost_2), each (B, n, n)."""
    out = _lib.teacher_volume(attn_list, None, plain_mean=True)
    return out.chunk(2, dim=0)


class VggtCostVolumes:
    """cost_1 / cost_2 of the VGGT teacher without the per-block attention maps.

    The reference collects ``attn`` (2 B, heads, n, n) from every global block (``vggt/layers/attention.py:73-84``,
    ``vggt/models/aggregator.py:259-260``), stacks and averages them (``:273``) and averages the heads
    (``src/finetune_timm_vggt.py:390-392``).  Here each block calls ``add_block(q, k)`` from inside its attention
    (in place of the ``return_attn`` branch) and the maps are reduced as they are produced::

        vols = VggtCostVolumes(num_blocks=len(aggregator.attn_indices), temperature=aggregator.temperature)
        ...  vols.add_block(q, k, self.scale)     # q, k: (B, heads, tokens, head_dim) after q_norm / k_norm / rope
        cost_1, cost_2 = vols.result()
    """

    def __init__(self, num_blocks, temperature=1.0, skip=5, round_bf16=True):
        self.weight = 1.0 / float(num_blocks)
        self.temperature, self.skip, self.round_bf16 = float(temperature), int(skip), bool(round_bf16)
        self.maps = None
        self.blocks = 0

    @torch.no_grad()
    def add_block(self, q, k, scale):
        # "q = q * self.scale" of the reference, in the tensors' own dtype (bf16 under the teacher's autocast)
        q_scaled = (q * scale).to(torch.bfloat16)
        self.maps = _lib.vggt_attn_accumulate(q_scaled, k.to(torch.bfloat16), self.temperature, self.weight, out=self.maps,
                                              skip=self.skip, round_bf16=self.round_bf16)
        self.blocks += 1

    def result(self, packed=False, eps=1e-8):
        """(cost_1, cost_2), each (B, n, n) fp32 -- or, with packed=True, each as the ``(fp16 volume, row statistics)``
        pair of ``gd3.ops.pack_teacher`` for ``gd3.ops.cost_volume_kl``."""
        if self.maps is None:
            raise _lib.Gd3Error('VggtCostVolumes.result() before any add_block()')
        if packed:
            from .. import ops
            return tuple(ops.pack_teacher(m, eps=eps) for m in self.maps)
        return self.maps
