"""Generate ``tests/golden/*.npz`` by running the LIVE reference functions.

Run in the build container only (``python -m oracle.gen_golden``): it needs the
reference checkout at ``/root/reference`` (read-only), which does not exist on the
GPU box.  The fixtures it writes are committed; tests never import the reference.

What is executed from the reference, unmodified:
  utils.losses.{kl_divergence_map, pairwise_logistic_ranking_loss, intra_depth_loss}
  utils.functions.{sigmoid, interpolate_features, extract_kp_depth,
                   get_patch_mask_from_kp_tensor, get_masked_patch_cost, filter_kp_by_conf}
  utils.model.DepthAwareFeatureFusion
  mast3r.fast_nn.{bruteforce_reciprocal_nns, fast_reciprocal_NNs, merge_corres}
  mast3r.losses.InfoNCE
  utils.functions.point_cloud_to_depth            (--depth-splat)
  vggt.layers.attention.Attention.custom_scaled_dot_product_attention   (--vggt-attn)
  dust3r.model.AsymmetricCroCo3DStereo.forward (tgt_attn_map)          (--teacher-volume)
``utils.functions`` imports kornia at module top (used only by an unrelated depth
filter); empty stub modules are registered for it.  The LightningModules cannot be
imported (timm / lightning / hydra absent), so the few lines of glue inside
``calculate_*_loss`` are re-typed here around the live functions, citing the lines.
"""

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get('GD3_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def import_reference():
    for name in ('kornia', 'kornia.filters', 'kornia.morphology'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import utils.losses as ref_losses
    import utils.functions as ref_functions
    import utils.model as ref_model
    import mast3r.fast_nn as ref_fast_nn
    return ref_losses, ref_functions, ref_model, ref_fast_nn


def ref_cost_loss(RL, RF, f1, f2, t12, t21, m1, m2, variant):
    """Glue of ``src/finetune_timm_mast3r.py:521-540`` / ``src/finetune_timm_vggt.py:510-533``."""
    a = F.normalize(f1[None], p=2, dim=-1)
    b = F.normalize(f2[None], p=2, dim=-1)
    c12 = torch.bmm(a, b.transpose(-1, -2))
    c21 = torch.bmm(b, a.transpose(-1, -2))
    if variant == 'mast3r':
        mt1 = RF.get_masked_patch_cost(t12.unsqueeze(0), m1, mask_patch_2=None)
        mt2 = RF.get_masked_patch_cost(t21.unsqueeze(0), m2, mask_patch_2=None)
        ms1 = RF.get_masked_patch_cost(c12, m1, mask_patch_2=None, use_softmax=True)
        ms2 = RF.get_masked_patch_cost(c21, m2, mask_patch_2=None, use_softmax=True)
    else:
        c12 = torch.nn.functional.softmax(c12, dim=-1)
        c21 = torch.nn.functional.softmax(c21, dim=-1)
        mt1 = RF.get_masked_patch_cost(t12.unsqueeze(0), m1, mask_patch_2=None)
        mt2 = RF.get_masked_patch_cost(t21.unsqueeze(0), m2, mask_patch_2=None)
        ms1 = RF.get_masked_patch_cost(c12, m1, mask_patch_2=None)
        ms2 = RF.get_masked_patch_cost(c21, m2, mask_patch_2=None)
    return (RL.kl_divergence_map(mt1, ms1) + RL.kl_divergence_map(mt2, ms2)) / 2


def ref_matching_loss(RF, d1, d2, p1, p2, variant, thr_neg=0.1, thr_pos=5e-3):
    """Glue of ``src/finetune_timm_mast3r.py:557-589``, ``..._vggt.py:543-574``, ``..._me.py:196-217``."""
    sig = RF.sigmoid
    desc_1, desc_2 = d1[None], d2[None]
    if variant == 'me':
        kp3d_dist = torch.cdist(p1[None], p2[None])
        sim = torch.bmm(desc_1, desc_2.transpose(-1, -2))
        pos_idxs = torch.nonzero(kp3d_dist < thr_pos, as_tuple=False)
        pos_sim = sim[pos_idxs[:, 0], pos_idxs[:, 1], pos_idxs[:, 2]]
        rpos = sig(pos_sim - 1., temp=0.01) + 1
        neg_mask = kp3d_dist[pos_idxs[:, 0], pos_idxs[:, 1]] > thr_neg
        rall = rpos + torch.sum(sig(sim[pos_idxs[:, 0], pos_idxs[:, 1]] - 1., temp=0.01) * neg_mask.float(), -1)
        ap1 = rpos / rall
        rpos = sig(1. - pos_sim, temp=0.01) + 1
        rall = rpos + torch.sum(
            sig(sim[pos_idxs[:, 0], pos_idxs[:, 1]] - pos_sim[:, None].repeat(1, sim.shape[-1]), temp=0.01)
            * neg_mask.float(), -1)
        ap2 = rpos / rall
        return torch.mean(1. - (ap1 + ap2) / 2)
    K = desc_1.size(1)
    pos_idxs = torch.stack([torch.zeros(K, dtype=torch.long), torch.arange(K), torch.arange(K)], dim=1)
    eye_mask = torch.eye(K).bool().unsqueeze(0)
    neg_mask = (torch.cdist(p1[None], p2[None]) > thr_neg) & ~eye_mask
    sim = torch.bmm(desc_1, desc_2.transpose(-1, -2))
    pos_sim = sim[pos_idxs[:, 0], pos_idxs[:, 1], pos_idxs[:, 2]]
    if variant == 'mast3r':
        rpos = sig(pos_sim - 1., temp=0.01) + 1
    else:
        rpos = sig(1. - pos_sim, temp=0.01) + 1
    rall = rpos + torch.sum(sig(sim[pos_idxs[:, 0], pos_idxs[:, 1]] - 1., temp=0.01)
                            * neg_mask[pos_idxs[:, 0], pos_idxs[:, 1]].float(), dim=-1)
    ap1 = rpos / rall
    rpos = sig(1. - pos_sim, temp=0.01) + 1
    rall = rpos + torch.sum(sig(sim[pos_idxs[:, 0], pos_idxs[:, 1]] - pos_sim[:, None], temp=0.01)
                            * neg_mask[pos_idxs[:, 0], pos_idxs[:, 1]].float(), dim=-1)
    ap2 = rpos / rall
    return torch.mean(1. - (ap1 + ap2) / 2)


def ref_sample_tokens(RF, tokens, ph, pw, kp, patch=14, normalize=False):
    """Glue of ``src/finetune_timm_mast3r.py:271-274`` / ``:307-313``."""
    P, N, C = tokens.shape
    res = tokens.reshape(P, ph, pw, -1).permute(0, 3, 1, 2).contiguous()
    out = RF.interpolate_features(res, kp, h=ph * patch, w=pw * patch, patch_size=patch, stride=patch,
                                  normalize=False).permute(0, 2, 1)
    if normalize:
        out = F.normalize(out, p=2, dim=-1)
    return out


def ref_depth_losses(RL, head, kf1, kf2, kd1, kd2):
    """Glue of ``src/finetune_timm_mast3r.py:489-499``."""
    pred = head(kf1 - kf2)
    depth_loss = F.l1_loss(pred, torch.tanh(kd1 - kd2).detach())
    l1 = RL.pairwise_logistic_ranking_loss(head, kf1, kd1, depth_threshold=0.05)
    l2 = RL.pairwise_logistic_ranking_loss(head, kf2, kd2, depth_threshold=0.05)
    return depth_loss, (l1 + l2) / 2


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def main():
    from oracle import synth
    RL, RF, RM, RN = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---------------- cost-volume KL ----------------
    kl = {}
    cases = [
        ('small_mast3r', 64, 32, 'mast3r', 'bernoulli', 11),
        ('small_vggt', 64, 32, 'vggt', 'bernoulli', 12),
        ('small_mast3r_allkept', 64, 32, 'mast3r', 'all', 13),
        ('small_mast3r_allmasked', 64, 32, 'mast3r', 'none', 14),
        ('small_vggt_allmasked', 64, 32, 'vggt', 'none', 15),
        ('odd_vggt', 49, 40, 'vggt', 'bernoulli', 16),          # N = 7^2, not tile aligned
        ('odd_mast3r', 169, 72, 'mast3r', 'bernoulli', 17),     # N = 13^2
        ('cfg1_mast3r', 256, 384, 'mast3r', 'bernoulli', 1000),
        ('mid_vggt', 400, 256, 'vggt', 'bernoulli', 18),
    ]
    for name, N, C, variant, mode, seed in cases:
        f1, f2 = synth.features(seed * 16, N, C)
        t12 = synth.teacher_volume(seed * 16 + 1, N, variant)
        t21 = synth.teacher_volume(seed * 16 + 2, N, variant)
        m1 = synth.patch_mask(seed * 16 + 3, N, mode=mode)
        m2 = synth.patch_mask(seed * 16 + 4, N, mode=mode)
        f1.requires_grad_(True)
        f2.requires_grad_(True)
        loss = ref_cost_loss(RL, RF, f1, f2, t12, t21, m1, m2, variant)
        if loss.requires_grad:
            g1, g2 = torch.autograd.grad(loss, [f1, f2], allow_unused=True)
            g1 = torch.zeros_like(f1) if g1 is None else g1
            g2 = torch.zeros_like(f2) if g2 is None else g2
        else:
            g1, g2 = torch.zeros_like(f1), torch.zeros_like(f2)
        kl[f'{name}/meta'] = np.array([N, C, seed, 0 if variant == 'mast3r' else 1,
                                       {'bernoulli': 0, 'all': 1, 'none': 2}[mode]], dtype=np.int64)
        kl[f'{name}/loss'] = _np(loss).astype(np.float64)
        kl[f'{name}/g1'] = _np(g1)
        kl[f'{name}/g2'] = _np(g2)
        print('kl', name, float(loss))
    np.savez_compressed(os.path.join(OUT, 'cost_kl.npz'), **kl)

    # ---------------- Smooth-AP ----------------
    ap = {}
    for name, K, C, variant, seed in [
        ('small_mast3r', 64, 32, 'mast3r', 21), ('small_vggt', 64, 32, 'vggt', 22),
        ('small_me', 64, 32, 'me', 23), ('k1_mast3r', 1, 32, 'mast3r', 24),
        ('odd_vggt', 77, 48, 'vggt', 25), ('cfg1_mast3r', 128, 384, 'mast3r', 1001),
        ('mid_me', 200, 96, 'me', 26),
    ]:
        N, ph, pw = 256, 16, 16
        g1, g2 = synth.ap_token_maps(seed * 16 + 5, N, C)
        kp1 = synth.keypoints(seed * 16 + 6, K, pw * 14, ph * 14)
        kp2 = kp1 + (synth.keypoints(seed * 16 + 7, K, 9, 9) - 4.0)   # nearby, so cosines stay high
        kp2[:, 0].clamp_(3, pw * 14 - 4)
        kp2[:, 1].clamp_(3, ph * 14 - 4)
        p1, p2 = synth.points3d(seed * 16 + 8, K)
        if variant == 'me':
            # several positives per row and some rows with none
            p2 = p1.clone()
            p2[::3] += 0.5
            p2[1::7] = p2[0::7][:p2[1::7].shape[0]]
            p2 = p2 + 1e-3 * torch.randn(K, 3, generator=synth._gen(seed))
        g1.requires_grad_(True)
        g2.requires_grad_(True)
        d1 = ref_sample_tokens(RF, g1[None], ph, pw, kp1[None], normalize=True)[0]
        d2 = ref_sample_tokens(RF, g2[None], ph, pw, kp2[None], normalize=True)[0]
        loss = ref_matching_loss(RF, d1, d2, p1, p2, variant)
        gg1, gg2 = torch.autograd.grad(loss, [g1, g2])
        d1l = d1.detach().clone().requires_grad_(True)
        d2l = d2.detach().clone().requires_grad_(True)
        l2 = ref_matching_loss(RF, d1l, d2l, p1, p2, variant)
        gd1, gd2 = torch.autograd.grad(l2, [d1l, d2l])
        ap[f'{name}/meta'] = np.array([K, C, seed, {'mast3r': 0, 'vggt': 1, 'me': 2}[variant]], dtype=np.int64)
        for k, v in dict(g1=g1, g2=g2, kp1=kp1, kp2=kp2, p1=p1, p2=p2, d1=d1, d2=d2, loss=loss.double(),
                         grad_g1=gg1, grad_g2=gg2, grad_d1=gd1, grad_d2=gd2).items():
            ap[f'{name}/{k}'] = _np(v)
        print('ap', name, float(loss), float(gg1.abs().max()))
    np.savez_compressed(os.path.join(OUT, 'smooth_ap.npz'), **ap)

    # ---------------- depth ranking / hinge / L1 ----------------
    rk = {}
    for name, K, D, seed, spread in [('small', 48, 64, 31, 1.0), ('novalid', 16, 32, 32, 0.0),
                                     ('cfg1', 128, 384, 1002, 1.0), ('k1', 1, 32, 33, 1.0),
                                     ('notanh', 40, 64, 34, 1.0)]:
        use_tanh = name != 'notanh'
        head = RM.DepthAwareFeatureFusion(D, use_tanh=use_tanh)
        hp = synth.head_params(seed * 16 + 11, D)
        synth.load_head(head, hp)
        g = synth._gen(seed * 16 + 12)
        kf1 = (0.5 * torch.randn(1, K, D, generator=g)).requires_grad_(True)
        kf2 = (0.5 * torch.randn(1, K, D, generator=g)).requires_grad_(True)
        kd1 = (1.0 + spread * synth.depths(seed * 16 + 9, K))[None]
        kd2 = (1.0 + spread * synth.depths(seed * 16 + 10, K))[None]
        params = list(head.fusion_layer.parameters())
        dl, il = ref_depth_losses(RL, head, kf1, kf2, kd1, kd2)
        hinge = RL.intra_depth_loss(head, kf1, kd1)
        outs = {}
        for tag, val in (('l1', dl), ('rank', il), ('hinge', hinge)):
            if val.requires_grad:
                grads = torch.autograd.grad(val, [kf1, kf2] + params, allow_unused=True, retain_graph=True)
            else:
                grads = [None] * (2 + len(params))
            names = ['kf1', 'kf2', 'W1', 'b1', 'gamma', 'beta', 'w2', 'b2']
            shapes = [kf1, kf2] + params
            outs[tag] = float(val)
            rk[f'{name}/{tag}/loss'] = np.float64(float(val))
            for n_, g_, s_ in zip(names, grads, shapes):
                rk[f'{name}/{tag}/grad_{n_}'] = _np(torch.zeros_like(s_) if g_ is None else g_)
        rk[f'{name}/meta'] = np.array([K, D, seed, int(use_tanh)], dtype=np.int64)
        for k, v in dict(kf1=kf1, kf2=kf2, kd1=kd1, kd2=kd2, **hp).items():
            rk[f'{name}/{k}'] = _np(v)
        print('rank', name, outs)
    np.savez_compressed(os.path.join(OUT, 'ranking.npz'), **rk)

    # ---------------- small helpers ----------------
    hp_ = {}
    g = synth._gen(41)
    fmap = torch.randn(2, 24, 9, 13, generator=g)
    pts = torch.stack([torch.rand(2, 50, generator=g) * (13 * 14 + 20) - 10,
                       torch.rand(2, 50, generator=g) * (9 * 14 + 20) - 10], dim=-1)   # incl. outside -> border clamp
    for nrm in (False, True):
        hp_[f'interp/out_norm{int(nrm)}'] = _np(RF.interpolate_features(fmap, pts, h=9 * 14, w=13 * 14, normalize=nrm))
    fm16 = torch.randn(1, 8, 20, 30, generator=g)
    pts16 = torch.stack([torch.rand(1, 40, generator=g) * 480, torch.rand(1, 40, generator=g) * 320], dim=-1)
    hp_['interp16/out'] = _np(RF.interpolate_features(fm16, pts16, h=320, w=480, normalize=False, patch_size=16, stride=16))
    hp_['interp/fmap'], hp_['interp/pts'] = _np(fmap), _np(pts)
    hp_['interp16/fmap'], hp_['interp16/pts'] = _np(fm16), _np(pts16)
    depth = torch.rand(40, 56, generator=g) * 4 + 0.5
    kpd = torch.stack([torch.randint(0, 56, (1, 30), generator=g), torch.randint(0, 40, (1, 30), generator=g)], dim=-1).float()
    hp_['kpdepth/depth'], hp_['kpdepth/kp'] = _np(depth), _np(kpd)
    hp_['kpdepth/out'] = _np(RF.extract_kp_depth(depth, kpd))
    kpm = torch.stack([torch.rand(60, generator=g) * 260 - 20, torch.rand(60, generator=g) * 200 - 20], dim=-1)
    hp_['kpmask/kp'] = _np(kpm)
    hp_['kpmask/out'] = _np(RF.get_patch_mask_from_kp_tensor(kpm, 168, 224, 14))
    hp_['kpmask/out_empty'] = _np(RF.get_patch_mask_from_kp_tensor(kpm - 1000, 168, 224, 14))
    xs = torch.linspace(-1.2, 1.2, 241)
    hp_['sigmoid/x'] = _np(xs)
    hp_['sigmoid/y_t001'] = _np(RF.sigmoid(xs, temp=0.01))
    hp_['sigmoid/y_t1'] = _np(RF.sigmoid(xs))
    cost = torch.rand(1, 12, 12, generator=g)
    mk1 = torch.rand(12, generator=g) < 0.5
    mk2 = torch.rand(12, generator=g) < 0.5
    hp_['mpc/cost'], hp_['mpc/m1'], hp_['mpc/m2'] = _np(cost), _np(mk1), _np(mk2)
    hp_['mpc/rownorm'] = _np(RF.get_masked_patch_cost(cost, mk1))
    hp_['mpc/softmax'] = _np(RF.get_masked_patch_cost(cost, mk1, use_softmax=True, temperature=0.5))
    hp_['mpc/rownorm_m2'] = _np(RF.get_masked_patch_cost(cost, mk1, mk2))
    tt = torch.softmax(3 * torch.randn(1, 12, 12, generator=g), -1)
    ss = torch.softmax(torch.randn(1, 12, 12, generator=g), -1)
    hp_['klmap/t'], hp_['klmap/s'] = _np(tt), _np(ss)
    hp_['klmap/out'] = _np(RL.kl_divergence_map(tt, ss)).astype(np.float64)
    conf = torch.rand(30, 40, generator=g) > 0.4
    kpc = torch.stack([torch.rand(1, 25, generator=g) * 38, torch.rand(1, 25, generator=g) * 28], dim=-1)
    fk, fi = RF.filter_kp_by_conf(kpc, conf)
    hp_['conf/mask'], hp_['conf/kp'], hp_['conf/idx'] = _np(conf), _np(kpc), _np(fi)
    # InfoNCE (upstream MASt3R softmax-CE correspondence loss)
    import mast3r.losses as ref_m3l
    dA = F.normalize(torch.randn(2, 33, 24, generator=g), dim=-1)
    dB = F.normalize(dA + 0.3 * torch.randn(2, 33, 24, generator=g), dim=-1)
    vm = torch.rand(2, 33, generator=g) < 0.8
    hp_['infonce/d1'], hp_['infonce/d2'], hp_['infonce/valid'] = _np(dA), _np(dB), _np(vm)
    for mode in ('all', 'proper', 'dual'):
        crit = ref_m3l.InfoNCE(mode=mode)
        hp_[f'infonce/{mode}'] = _np(crit(dA, dB, valid_matches=vm)).astype(np.float64)
    np.savez_compressed(os.path.join(OUT, 'helpers.npz'), **hp_)
    print('helpers done')

    # ---------------- fast_nn ----------------
    nn = {}
    A = synth.nn_exact_set(51, 512)
    B = synth.nn_exact_set(52, 384)
    nn['exact/A'], nn['exact/B'] = _np(A), _np(B)
    for tag, kw in (('dot', dict(dist='dot')), ('dot_blk', dict(dist='dot', block_size=128)),
                    ('l2', dict(dist='l2')), ('l2_blk', dict(dist='l2', block_size=100))):
        a, b = RN.bruteforce_reciprocal_nns(A, B, device='cpu', **kw)
        nn[f'exact/{tag}/nnA'], nn[f'exact/{tag}/nnB'] = a, b
    a, b = RN.bruteforce_reciprocal_nns(A.numpy(), B.numpy(), device='cpu', dist='dot')
    assert (a == nn['exact/dot/nnA']).all() and (b == nn['exact/dot/nnB']).all()
    # dense descriptor maps -> fast_reciprocal_NNs (blocked brute-force path, as the trainer calls it)
    # descriptors quantised to k/8 so that every dot product is exact and the result
    # does not depend on the matmul's summation order
    d1, d2 = synth.nn_desc_maps(53, 48, 64)
    d1 = torch.round(d1 * 16) / 8
    d2 = torch.round(d2 * 16) / 8
    nn['maps/d1'], nn['maps/d2'] = _np(d1), _np(d2)
    for tag, kw in (('s8', dict(subsample_or_initxy1=8)), ('s4', dict(subsample_or_initxy1=4))):
        xy1, xy2 = RN.fast_reciprocal_NNs(d1, d2, device='cpu', dist='dot', block_size=2 ** 10, **kw)
        nn[f'maps/{tag}/xy1'], nn[f'maps/{tag}/xy2'] = xy1, xy2
        print('fast_nn', tag, xy1.shape)
    i1, i2 = RN.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=8, ret_xy=False, device='cpu', dist='dot', block_size=2 ** 10)
    nn['maps/s8_idx/i1'], nn['maps/s8_idx/i2'] = i1, i2
    xs_, ys_ = np.mgrid[2:64:7, 3:48:5].reshape(2, -1)
    xy1, xy2 = RN.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=(xs_, ys_), pixel_tol=3, device='cpu', dist='dot', block_size=2 ** 10)
    nn['maps/seeds/x'], nn['maps/seeds/y'] = xs_, ys_
    nn['maps/seeds_tol3/xy1'], nn['maps/seeds_tol3/xy2'] = xy1, xy2
    xy1, xy2, basin = RN.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=8, ret_basin=True, device='cpu', dist='dot', block_size=2 ** 10)
    nn['maps/basin/xy1'], nn['maps/basin/xy2'], nn['maps/basin/basin'] = xy1, xy2, basin
    # merge_corres KAT
    gi = np.random.RandomState(7)
    m1 = gi.randint(0, 48 * 64, 300).astype(np.int32)
    m2 = gi.randint(0, 48 * 64, 300).astype(np.int32)
    m1[100:150], m2[100:150] = m1[:50], m2[:50]
    nn['merge/i1'], nn['merge/i2'] = m1, m2
    x1, x2 = RN.merge_corres(m1, m2, (48, 64), (48, 64))
    nn['merge/xy1'], nn['merge/xy2'] = x1, x2
    j1, j2, jdx = RN.merge_corres(m1, m2, ret_xy=False, ret_index=True)
    nn['merge/j1'], nn['merge/j2'], nn['merge/jdx'] = j1, j2, jdx
    np.savez_compressed(os.path.join(OUT, 'fast_nn.npz'), **nn)
    print('fast_nn done')


def fast_nn_extra():
    """``extract_correspondences_nonsym`` (``mast3r/fast_nn.py:191-223``) from the live reference ->
    ``tests/golden/fast_nn_extra.npz``.  Separate file so that the other golden files stay byte-identical."""
    from oracle import synth
    _, _, _, RN = import_reference()
    out = {}
    d1, d2 = synth.nn_desc_maps(61, 40, 56)
    d1 = torch.round(d1 * 16) / 8
    d2 = torch.round(d2 * 16) / 8
    g = torch.Generator().manual_seed(62)
    cA = torch.rand(40, 56, generator=g) + 1.0
    cB = torch.rand(40, 56, generator=g) + 1.0
    out['d1'], out['d2'], out['cA'], out['cB'] = _np(d1), _np(d2), _np(cA), _np(cB)
    for tag, tol in (('tol0', 0), ('tol2', 2)):
        xy1, xy2, conf = RN.extract_correspondences_nonsym(d1, d2, cA.numpy(), cB.numpy(), subsample=8, device='cpu',
                                                           pixel_tol=tol)
        out[f'{tag}/xy1'], out[f'{tag}/xy2'], out[f'{tag}/conf'] = _np(xy1), _np(xy2), _np(conf)
        print('extract_correspondences_nonsym', tag, tuple(xy1.shape))
    np.savez_compressed(os.path.join(OUT, 'fast_nn_extra.npz'), **out)


def depth_splat():
    """``point_cloud_to_depth`` (``utils/functions.py:218-260``) from the live reference -> ``tests/golden/depth_splat.npz``."""
    from oracle import synth
    _, RF, _, _ = import_reference()
    out = {}
    for name, (pts, K, w, h) in synth.depth_splat_cases().items():
        ref = RF.point_cloud_to_depth(pts, K, w, h, 'cpu')
        out[f'{name}/pts'], out[f'{name}/K'], out[f'{name}/wh'] = _np(pts), _np(K), np.array([w, h])
        out[f'{name}/depth'] = _np(ref)
        print('point_cloud_to_depth', name, tuple(ref.shape), 'filled', int((ref > 0).sum()))
    np.savez_compressed(os.path.join(OUT, 'depth_splat.npz'), **out)


def vggt_attn():
    """The ``return_attn`` branch of the live ``vggt.layers.attention.Attention`` (``:73-84``) ->
    ``tests/golden/vggt_attn.npz``.  Two runs per case on the same bf16-representable q / k: fp32 tensors (every op
    fp32) and bf16 tensors without autocast (bf16 scores, bf16 scores / temperature, softmax evaluated in fp32 and
    rounded to bf16 on output); CPU autocast is not used because its op lists differ from CUDA's (softmax stays bf16)."""
    from oracle import synth
    import_reference()
    from vggt.layers.attention import Attention
    out = {}
    for name, (B, heads, n, temp) in {'small': (1, 3, 37, 1.0), 'ragged_t3': (2, 2, 45, 3.0), 'blocks': (1, 4, 64, 2.0)}.items():
        att = Attention(dim=64 * heads, num_heads=heads)
        blocks32 = []
        for blk in range(3 if name == 'blocks' else 1):
            q, k = synth.vggt_qk(900 + 10 * len(out) + blk, B, heads, n)
            v = torch.zeros_like(q)
            with torch.no_grad():
                _, a32 = att.custom_scaled_dot_product_attention(q.float(), k.float(), v.float(), return_attn=True,
                                                                 temperature=temp)
                _, a16 = att.custom_scaled_dot_product_attention(q, k, v, return_attn=True, temperature=temp)
            assert a32.dtype == torch.float32 and a16.dtype == torch.bfloat16
            out[f'{name}/q{blk}'], out[f'{name}/k{blk}'] = _np(q.float()), _np(k.float())
            out[f'{name}/attn_fp32_{blk}'], out[f'{name}/attn_bf16_{blk}'] = _np(a32), _np(a16.float())
            blocks32.append(a32)
        # vggt/models/aggregator.py:273 and src/finetune_timm_vggt.py:390-392 on the live maps
        attn_mean = torch.mean(torch.stack(blocks32), dim=0)
        cost_1, cost_2 = attn_mean.chunk(2, dim=0)
        out[f'{name}/cost_1'], out[f'{name}/cost_2'] = _np(cost_1.mean(dim=1)), _np(cost_2.mean(dim=1))
        out[f'{name}/meta'] = np.array([B, heads, n, len(blocks32)], dtype=np.int64)
        out[f'{name}/temperature'] = np.array(temp, dtype=np.float32)
        out[f'{name}/scale'] = np.array(att.scale, dtype=np.float32)
        print('vggt attention', name, tuple(a32.shape))
    np.savez_compressed(os.path.join(OUT, 'vggt_attn.npz'), **out)


def teacher_volume():
    """``tgt_attn_map`` of the live MASt3R teacher class (``dust3r/dust3r/model.py:346-363``) ->
    ``tests/golden/teacher_volume.npz``.  A small random-weight ``AsymmetricCroCo3DStereo`` (3 decoder layers, 3 heads,
    4 x 6 patches) runs its unmodified ``forward`` on CPU; the per-layer cross-attention logits that ``_decoder``
    returns are recorded on the way (they are the inputs of the block), with the cross-attention query / key
    projections scaled up so that the softmax rows are peaked.  ``timm`` is imported by croco's blocks.py but never used:
    an empty stub module is registered for it."""
    sys.modules.setdefault('timm', types.ModuleType('timm'))
    for pth in (REF, os.path.join(REF, 'dust3r')):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import dust3r.model as dust3r_model
    out = {}
    for name, (B, reciprocity, temperature) in {'recip': (2, True, 3.0), 'recip_t1': (1, True, 1.0),
                                                'plain': (2, False, 3.0)}.items():
        torch.manual_seed(31 + len(out))
        net = dust3r_model.AsymmetricCroCo3DStereo(
            img_size=(64, 96), patch_size=16, enc_embed_dim=64, enc_depth=1, enc_num_heads=2, dec_embed_dim=48,
            dec_depth=3, dec_num_heads=3, pos_embed='RoPE100', head_type='linear', output_mode='pts3d',
            landscape_only=False, temperature=temperature)
        net.eval()
        net.reciprocity = reciprocity
        with torch.no_grad():
            for n_, prm in net.named_parameters():
                if 'cross_attn.projq.weight' in n_ or 'cross_attn.projk.weight' in n_:
                    prm.mul_(6.0)
        seen = {}
        inner = net._decoder

        def recording_decoder(*a, _inner=inner, _seen=seen, **k):
            res = _inner(*a, **k)
            _seen['tgt'], _seen['src'] = [c.clone() for c in res[1]], [c.clone() for c in res[2]]
            return res
        net._decoder = recording_decoder
        views = [dict(img=torch.randn(B, 3, 64, 96), true_shape=torch.tensor([[64, 96]] * B), instance=[str(i)] * B)
                 for i in range(2)]
        with torch.no_grad():
            _, res2 = net(*views)
        out[f'{name}/tgt'] = _np(torch.stack(seen['tgt']))          # (L, B, heads, N, N)
        out[f'{name}/src'] = _np(torch.stack(seen['src']))
        out[f'{name}/tgt_attn_map'] = _np(res2['tgt_attn_map'])
        out[f'{name}/temperature'] = np.array(temperature, dtype=np.float32)
        out[f'{name}/reciprocity'] = np.array(int(reciprocity))
        print('teacher_volume', name, tuple(res2['tgt_attn_map'].shape), 'row max', float(res2['tgt_attn_map'].max(-1).values.mean()))
    np.savez_compressed(os.path.join(OUT, 'teacher_volume.npz'), **out)


if __name__ == '__main__':
    if '--fast-nn-extra' in sys.argv:
        fast_nn_extra()
    elif '--teacher-volume' in sys.argv:
        teacher_volume()
    elif '--vggt-attn' in sys.argv:
        vggt_attn()
    elif '--depth-splat' in sys.argv:
        depth_splat()
    else:
        main()
