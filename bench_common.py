"""Shared by bench.py, __graft_entry__.smoke() and the step-level tests: synthetic batches (SURVEY.md
section 8-d) and the CPU oracle evaluation of one distillation step.

This module imports ``oracle`` -- it is bench / test infrastructure, not product code.
"""
import torch

from oracle import bodies, losses as olosses, synth

WEIGHTS = {
    'mast3r': dict(ap=1.0, depth=0.0, intra=1.0, kl=1.0),   # src/finetune_timm_mast3r.py:79-84
    'vggt': dict(ap=1.0, depth=1.0, intra=1.0, kl=1.0),     # src/finetune_timm_vggt.py:86-89
}
CFG_IDS = {'cfg1': 1, 'cfg2': 2, 'cfg4': 4}


def make_batch(cfg, cfg_id=2, pair0=0, device='cpu', pairs=None):
    """Seeded CPU batch of ``pairs`` (default cfg['P']) pairs: dict of stacked fp32 tensors + head params."""
    P = cfg['P'] if pairs is None else pairs
    N, C, K, grid, variant = cfg['N'], cfg['C'], cfg['K'], cfg['grid'], cfg['variant']
    items = [synth.pair_inputs(cfg_id, pair0 + p, N, C, K, grid, variant) for p in range(P)]

    def st(name):
        return torch.stack([getattr(it, name) for it in items])
    batch = dict(f1=st('f1'), f2=st('f2'), t12=st('t12'), t21=st('t21'), m1=st('m1'), m2=st('m2'),
                 g1=st('g1'), g2=st('g2'), kp1=st('kp1'), kp2=st('kp2'), p3d1=st('p1'), p3d2=st('p2'),
                 dep1=st('d1'), dep2=st('d2'))
    # keypoints of view 2 close to those of view 1, so sampled descriptors correlate (non-trivial AP gradients)
    ph, pw = grid
    jitter = torch.stack([synth.keypoints(977 * cfg_id + pair0 + p, K, 9, 9) - 4.0 for p in range(P)])
    kp2 = batch['kp1'] + jitter
    kp2[..., 0].clamp_(3, pw * 14 - 4)
    kp2[..., 1].clamp_(3, ph * 14 - 4)
    batch['kp2'] = kp2
    hp = synth.head_params(4242 + cfg_id, C)
    batch['head'] = dict(hp, use_tanh=True, ln_eps=1e-5)
    if device != 'cpu':
        batch = to_device(batch, device)
    return batch


def to_device(batch, device, feature_dtype=None, non_blocking=False):
    out = {}
    for k, v in batch.items():
        if k == 'head':
            out[k] = {n: (t.to(device, non_blocking=non_blocking) if torch.is_tensor(t) else t) for n, t in v.items()}
        elif torch.is_tensor(v):
            t = v.to(device, non_blocking=non_blocking)
            if feature_dtype is not None and k in ('f1', 'f2', 'g1', 'g2', 'h1', 'h2'):
                t = t.to(feature_dtype)
            out[k] = t
        else:
            out[k] = v
    return out


def oracle_head(batch, C):
    head = olosses.DepthHead(C, use_tanh=batch['head'].get('use_tanh', True))
    synth.load_head(head, batch['head'])
    return head


def oracle_step(batch, cfg, backward=True, pairs=None):
    """The reference's per-pair loss flow (B = 1 per call) on CPU fp32 through the oracle.

    Follows training_step steps 3-6 (src/finetune_timm_mast3r.py:636-653): depth losses, KL, Smooth-AP,
    weighted sum; total = mean over pairs.  Returns per-pair losses and gradients w.r.t. f1, f2, g1, g2
    and the head parameters (packed like the CUDA path).
    """
    variant, (ph, pw) = cfg['variant'], cfg['grid']
    w = WEIGHTS[variant]
    P = batch['f1'].shape[0] if pairs is None else pairs
    C = batch['f1'].shape[2]
    head = oracle_head(batch, batch['head']['W1'].shape[1])
    leaves = {k: batch[k][:P].clone().float().requires_grad_(backward) for k in ('f1', 'f2', 'g1', 'g2')}
    # optional separate maps for the depth-head features, (P, N, C) or an (L, P, N, C) stack of ViT blocks whose
    # samples are averaged (get_intermediate_feature, src/finetune_timm_mast3r.py:271-277)
    for k in ('h1', 'h2'):
        if k in batch:
            t = batch[k]
            leaves[k] = (t[:, :P] if t.dim() == 4 else t[:P]).clone().float().requires_grad_(backward)

    def depth_feats(name, fallback, p, kp):
        if name not in leaves:
            return bodies.sample_tokens(fallback, ph, pw, kp, normalize=False)
        t = leaves[name]
        if t.dim() == 3:
            return bodies.sample_tokens(t[p:p + 1], ph, pw, kp, normalize=False)
        return torch.stack([bodies.sample_tokens(t[l, p:p + 1], ph, pw, kp, normalize=False)
                            for l in range(t.shape[0])]).mean(dim=0)
    res = {k: [] for k in ('kl', 'ap', 'rank', 'l1')}
    total = 0.0
    for p in range(P):
        kl = bodies.cost_volume_kl(leaves['f1'][p], leaves['f2'][p], batch['t12'][p], batch['t21'][p],
                                   batch['m1'][p], batch['m2'][p], variant)
        g1, g2 = leaves['g1'][p:p + 1], leaves['g2'][p:p + 1]
        kp1, kp2 = batch['kp1'][p:p + 1], batch['kp2'][p:p + 1]
        d1 = bodies.sample_tokens(g1, ph, pw, kp1, normalize=True)[0]
        d2 = bodies.sample_tokens(g2, ph, pw, kp2, normalize=True)[0]
        ap = bodies.smooth_ap(d1, d2, batch['p3d1'][p], batch['p3d2'][p], variant)
        kf1 = depth_feats('h1', g1, p, kp1)
        kf2 = depth_feats('h2', g2, p, kp2)
        l1, rank = bodies.depth_losses(head, kf1, kf2, batch['dep1'][p:p + 1], batch['dep2'][p:p + 1])
        total = total + (w['ap'] * ap + w['depth'] * l1 + w['intra'] * rank + w['kl'] * kl) / P
        for k, v in (('kl', kl), ('ap', ap), ('rank', rank), ('l1', l1)):
            res[k].append(v.detach())
    out = {k: torch.stack(v) for k, v in res.items()}
    out['total'] = total.detach()
    if backward:
        total.backward()
        fl = head.fusion_layer
        ps = [fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight, fl[3].bias]
        packed = torch.cat([(torch.zeros_like(q) if q.grad is None else q.grad).reshape(-1) for q in ps])
        out['grads'] = dict(head=packed, **{k: v.grad for k, v in leaves.items()})
    return out
