"""Oracle restatement of the teachers' cost-volume post-processing (MASt3R) and cross-view attention maps (VGGT).

Test infrastructure only (see ``oracle/__init__.py``).  Plain PyTorch on CPU.

``teacher_volume`` is pinned: ``oracle/gen_golden.py --teacher-volume`` instantiates a small random-weight
``AsymmetricCroCo3DStereo`` from the reference checkout, runs its unmodified ``forward`` on CPU and records the per-layer
logits together with the ``tgt_attn_map`` it returns (``tests/golden/teacher_volume.npz``).
``vggt_block_attention`` is pinned: the reference's ``Attention`` class imports and runs here on CPU.
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


def vggt_block_attention(q, k, scale, temperature=1.0, skip=5):
    """Cross-view attention maps of one VGGT global block.  Follows the ``return_attn`` branch of
    ``vggt/layers/attention.py:60,73-84``: q is scaled, the patch tokens of view 1 (rows ``skip .. N//2``) attend to
    those of view 2 (rows ``N//2 + skip ..``) and vice versa, ``softmax(scores / temperature)`` per head; the two
    directions are concatenated on the batch axis.

    Pinned against the live reference class (``tests/golden/vggt_attn.npz``).  With bf16 q / k the score matmul and the
    division return bf16 and the softmax runs in fp32 -- what CUDA bf16 autocast does in the reference
    (``src/finetune_timm_vggt.py:359``); with fp32 q / k everything is fp32.

    q, k: (B, heads, N, head_dim).  Returns (2 B, heads, n, n) fp32, n = N // 2 - skip.
    """
    N = q.shape[-2]
    q = q * scale
    view1, view2 = slice(skip, N // 2), slice(N // 2 + skip, None)
    s12 = torch.matmul(q[..., view1, :], k[..., view2, :].transpose(-2, -1))
    s21 = torch.matmul(q[..., view2, :], k[..., view1, :].transpose(-2, -1))
    a12 = torch.softmax((s12 / temperature).float(), dim=-1)
    a21 = torch.softmax((s21 / temperature).float(), dim=-1)
    return torch.cat([a12, a21], dim=0)


def vggt_cost_volumes(block_maps):
    """Mean over the collected blocks (``vggt/models/aggregator.py:273``), split into the two directions and mean
    over heads (``src/finetune_timm_vggt.py:390-392``).  block_maps: list of (2 B, heads, n, n) -> (cost_1, cost_2)."""
    attn = torch.mean(torch.stack(block_maps), dim=0)
    cost_1, cost_2 = attn.chunk(2, dim=0)
    return cost_1.mean(dim=1), cost_2.mean(dim=1)
