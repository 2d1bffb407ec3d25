"""Oracle restatement of the MASt3R teacher's cost-volume post-processing.

Test infrastructure only (see ``oracle/__init__.py``).  Plain PyTorch on CPU.

**Parity unpinned**: the block sits inside ``AsymmetricCroCo3DStereo.forward`` and cannot be called without the teacher model and its weights, so no golden vector could be
produced by the live reference; the restatement re-types those lines around the same torch ops.
"""
import torch


def teacher_volume(tgt_camap, src_camap, temperature=3.0, reciprocity=True):
    """Follows ``dust3r/dust3r/model.py:346-363`` line by line.

    tgt_camap / src_camap: lists (one entry per decoder layer) of pre-softmax cross-attention
    logits (B, heads, N, N) as returned by ``dust3r/croco/models/blocks.py:163-164``.
    Returns tgt_attn_map (B, N, N).
    """
    if reciprocity:
        tgt = [a.mean(dim=1).detach() for a in tgt_camap]
        src = [a.mean(dim=1).detach() for a in src_camap]
        tgt = [(t + s.transpose(-1, -2)) / 2 for t, s in zip(tgt, src)]
        tgt = [(c / temperature).softmax(dim=-1) for c in tgt]
        for i in range(len(tgt)):
            tgt[i][:, :, 0] = tgt[i].min()
    else:
        tgt = [a.mean(dim=1).detach() for a in tgt_camap]
        for i in range(len(tgt)):
            tgt[i][:, :, 0] = tgt[i].min()
    return torch.stack(tgt, dim=1).mean(dim=1)


def vggt_cost_volumes(attn_list):
    """Follows ``vggt/models/aggregator.py:273`` and ``src/finetune_timm_vggt.py:390-392``: mean over the
    global blocks, split into the two directions, mean over heads.  attn_list: list of (2B, heads, n, n)."""
    attn = torch.mean(torch.stack(attn_list), dim=0)
    cost_1, cost_2 = attn.chunk(2, dim=0)
    return cost_1.mean(dim=1), cost_2.mean(dim=1)
