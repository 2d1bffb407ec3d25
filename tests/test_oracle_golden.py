"""Pin the CPU oracle against golden vectors produced by the LIVE reference functions
(``oracle/gen_golden.py``).  CPU only; never imports the reference."""
import numpy as np
import pytest
import torch

from oracle import bodies, fast_nn, functions, losses, synth
from helpers import assert_grad_close, rel_err

T = torch.as_tensor

KL_CASES = ['small_mast3r', 'small_vggt', 'small_mast3r_allkept', 'small_mast3r_allmasked',
            'small_vggt_allmasked', 'odd_vggt', 'odd_mast3r', 'cfg1_mast3r', 'mid_vggt']


def kl_inputs(meta):
    N, C, seed, var, mode = [int(v) for v in meta]
    variant = 'mast3r' if var == 0 else 'vggt'
    mode = ['bernoulli', 'all', 'none'][mode]
    f1, f2 = synth.features(seed * 16, N, C)
    t12 = synth.teacher_volume(seed * 16 + 1, N, variant)
    t21 = synth.teacher_volume(seed * 16 + 2, N, variant)
    m1 = synth.patch_mask(seed * 16 + 3, N, mode=mode)
    m2 = synth.patch_mask(seed * 16 + 4, N, mode=mode)
    return f1, f2, t12, t21, m1, m2, variant


@pytest.mark.parametrize('case', KL_CASES)
def test_cost_kl_oracle(golden, case):
    g = golden('cost_kl.npz')
    f1, f2, t12, t21, m1, m2, variant = kl_inputs(g[f'{case}/meta'])
    f1.requires_grad_(True)
    f2.requires_grad_(True)
    loss = bodies.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant)
    want = float(g[f'{case}/loss'])
    assert abs(float(loss) - want) <= 1e-5 * max(1.0, abs(want))
    if loss.requires_grad and (m1.any() or m2.any()):
        g1, g2 = torch.autograd.grad(loss, [f1, f2])
        assert_grad_close(g1, T(g[f'{case}/g1']), cos_min=0.99999, name='g1', norm_rtol=1e-3)
        assert_grad_close(g2, T(g[f'{case}/g2']), cos_min=0.99999, name='g2', norm_rtol=1e-3)
    else:
        assert float(np.abs(g[f'{case}/g1']).max()) == 0.0


AP_CASES = ['small_mast3r', 'small_vggt', 'small_me', 'k1_mast3r', 'odd_vggt', 'cfg1_mast3r', 'mid_me']


@pytest.mark.parametrize('case', AP_CASES)
def test_smooth_ap_oracle(golden, case):
    g = golden('smooth_ap.npz')
    K, C, seed, var = [int(v) for v in g[f'{case}/meta']]
    variant = ['mast3r', 'vggt', 'me'][var]
    g1 = T(g[f'{case}/g1']).requires_grad_(True)
    g2 = T(g[f'{case}/g2']).requires_grad_(True)
    kp1, kp2 = T(g[f'{case}/kp1']), T(g[f'{case}/kp2'])
    p1, p2 = T(g[f'{case}/p1']), T(g[f'{case}/p2'])
    d1 = bodies.sample_tokens(g1[None], 16, 16, kp1[None], normalize=True)[0]
    d2 = bodies.sample_tokens(g2[None], 16, 16, kp2[None], normalize=True)[0]
    np.testing.assert_allclose(d1.detach().numpy(), g[f'{case}/d1'], rtol=0, atol=1e-6)
    loss = bodies.smooth_ap(d1, d2, p1, p2, variant)
    assert rel_err(loss, g[f'{case}/loss']) <= 1e-5
    gg1, gg2 = torch.autograd.grad(loss, [g1, g2])
    assert_grad_close(gg1, T(g[f'{case}/grad_g1']), cos_min=0.9999, name='grad_g1', norm_rtol=1e-2)
    assert_grad_close(gg2, T(g[f'{case}/grad_g2']), cos_min=0.9999, name='grad_g2', norm_rtol=1e-2)


RANK_CASES = ['small', 'novalid', 'cfg1', 'k1', 'notanh']


def load_head_case(g, case):
    K, D, seed, use_tanh = [int(v) for v in g[f'{case}/meta']]
    head = losses.DepthHead(D, use_tanh=bool(use_tanh))
    synth.load_head(head, {k: T(g[f'{case}/{k}']) for k in ('W1', 'b1', 'gamma', 'beta', 'w2', 'b2')})
    kf1 = T(g[f'{case}/kf1']).requires_grad_(True)
    kf2 = T(g[f'{case}/kf2']).requires_grad_(True)
    return head, kf1, kf2, T(g[f'{case}/kd1']), T(g[f'{case}/kd2'])


@pytest.mark.parametrize('case', RANK_CASES)
def test_depth_losses_oracle(golden, case):
    g = golden('ranking.npz')
    head, kf1, kf2, kd1, kd2 = load_head_case(g, case)
    params = list(head.fusion_layer.parameters())
    names = ['kf1', 'kf2', 'W1', 'b1', 'gamma', 'beta', 'w2', 'b2']
    l1, rank = bodies.depth_losses(head, kf1, kf2, kd1, kd2)
    hinge = losses.intra_depth_loss(head, kf1, kd1)
    for tag, val in (('l1', l1), ('rank', rank), ('hinge', hinge)):
        want = float(g[f'{case}/{tag}/loss'])
        assert abs(float(val) - want) <= 2e-6 * max(1.0, abs(want)), tag
        if not val.requires_grad:
            assert want == 0.0
            continue
        grads = torch.autograd.grad(val, [kf1, kf2] + params, allow_unused=True, retain_graph=True)
        for n_, got in zip(names, grads):
            ref = T(g[f'{case}/{tag}/grad_{n_}'])
            got = torch.zeros_like(ref) if got is None else got
            assert_grad_close(got, ref, cos_min=0.9999, name=f'{tag}/{n_}', norm_rtol=1e-2)


def test_helpers_oracle(golden):
    g = golden('helpers.npz')
    fmap, pts = T(g['interp/fmap']), T(g['interp/pts'])
    for nrm in (False, True):
        out = functions.interpolate_features(fmap, pts, h=9 * 14, w=13 * 14, normalize=nrm)
        np.testing.assert_allclose(out.numpy(), g[f'interp/out_norm{int(nrm)}'], rtol=0, atol=1e-6)
    out = functions.interpolate_features(T(g['interp16/fmap']), T(g['interp16/pts']), h=320, w=480,
                                         normalize=False, patch_size=16, stride=16)
    np.testing.assert_allclose(out.numpy(), g['interp16/out'], rtol=0, atol=1e-6)
    out = functions.extract_kp_depth(T(g['kpdepth/depth']), T(g['kpdepth/kp']))
    np.testing.assert_allclose(out.numpy(), g['kpdepth/out'], rtol=0, atol=1e-6)
    kp = T(g['kpmask/kp'])
    assert (functions.get_patch_mask_from_kp_tensor(kp, 168, 224, 14).numpy() == g['kpmask/out']).all()
    assert (functions.get_patch_mask_from_kp_tensor(kp - 1000, 168, 224, 14).numpy() == g['kpmask/out_empty']).all()
    xs = T(g['sigmoid/x'])
    np.testing.assert_allclose(functions.sigmoid(xs, 0.01).numpy(), g['sigmoid/y_t001'], rtol=1e-6, atol=0)
    np.testing.assert_allclose(functions.sigmoid(xs).numpy(), g['sigmoid/y_t1'], rtol=1e-6, atol=0)
    cost, m1, m2 = T(g['mpc/cost']), T(g['mpc/m1']), T(g['mpc/m2'])
    np.testing.assert_allclose(functions.get_masked_patch_cost(cost, m1).numpy(), g['mpc/rownorm'], rtol=1e-6)
    np.testing.assert_allclose(functions.get_masked_patch_cost(cost, m1, use_softmax=True, temperature=0.5).numpy(),
                               g['mpc/softmax'], rtol=1e-6)
    np.testing.assert_allclose(functions.get_masked_patch_cost(cost, m1, m2).numpy(), g['mpc/rownorm_m2'], rtol=1e-6)
    assert rel_err(losses.kl_divergence_map(T(g['klmap/t']), T(g['klmap/s'])), g['klmap/out']) < 1e-6
    _, idx = functions.filter_kp_by_conf(T(g['conf/kp']), T(g['conf/mask']))
    assert (idx.numpy() == g['conf/idx']).all()
    for mode in ('all', 'proper', 'dual'):
        got = losses.infonce(T(g['infonce/d1']), T(g['infonce/d2']), T(g['infonce/valid']), mode=mode)
        assert rel_err(got, g[f'infonce/{mode}']) < 1e-5


def test_fast_nn_oracle(golden):
    g = golden('fast_nn.npz')
    A, B = T(g['exact/A']), T(g['exact/B'])
    for tag, kw in (('dot', dict(dist='dot')), ('dot_blk', dict(dist='dot', block_size=128)),
                    ('l2', dict(dist='l2')), ('l2_blk', dict(dist='l2', block_size=100))):
        a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cpu', **kw)
        assert a.dtype == np.int64 and b.dtype == np.int64
        assert (a == g[f'exact/{tag}/nnA']).all(), tag
        assert (b == g[f'exact/{tag}/nnB']).all(), tag
    # blocked == single shot on exact inputs (lowest-index ties in both)
    assert (g['exact/dot/nnA'] == g['exact/dot_blk/nnA']).all()
    assert (g['exact/dot/nnB'] == g['exact/dot_blk/nnB']).all()
    with pytest.raises(ValueError):
        fast_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='cosine')
    d1, d2 = T(g['maps/d1']), T(g['maps/d2'])
    for tag, kw in (('s8', dict(subsample_or_initxy1=8)), ('s4', dict(subsample_or_initxy1=4))):
        xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, device='cpu', dist='dot', block_size=2 ** 10, **kw)
        assert (xy1 == g[f'maps/{tag}/xy1']).all() and (xy2 == g[f'maps/{tag}/xy2']).all()
        assert xy1.dtype == g[f'maps/{tag}/xy1'].dtype
    i1, i2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=8, ret_xy=False, device='cpu',
                                         dist='dot', block_size=2 ** 10)
    assert (i1 == g['maps/s8_idx/i1']).all() and (i2 == g['maps/s8_idx/i2']).all()
    xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=(g['maps/seeds/x'], g['maps/seeds/y']),
                                           pixel_tol=3, device='cpu', dist='dot', block_size=2 ** 10)
    assert (xy1 == g['maps/seeds_tol3/xy1']).all() and (xy2 == g['maps/seeds_tol3/xy2']).all()
    xy1, xy2, basin = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=8, ret_basin=True, device='cpu',
                                                  dist='dot', block_size=2 ** 10)
    assert (xy1 == g['maps/basin/xy1']).all() and (basin == g['maps/basin/basin']).all()
    x1, x2 = fast_nn.merge_corres(g['merge/i1'], g['merge/i2'], (48, 64), (48, 64))
    assert (x1 == g['merge/xy1']).all() and (x2 == g['merge/xy2']).all()
    j1, j2, jdx = fast_nn.merge_corres(g['merge/i1'], g['merge/i2'], ret_xy=False, ret_index=True)
    assert (j1 == g['merge/j1']).all() and (j2 == g['merge/j2']).all() and (jdx == g['merge/jdx']).all()
    with pytest.raises(AssertionError):
        fast_nn.merge_corres(g['merge/i1'].astype(np.int64), g['merge/i2'].astype(np.int64))
    # KDTree branch of the CPU path agrees with brute-force l2 on generic inputs
    p1 = torch.rand(12, 16, 3, generator=synth._gen(3))
    p2 = p1 + 0.01 * torch.rand(12, 16, 3, generator=synth._gen(4))
    a1, a2 = fast_nn.fast_reciprocal_NNs(p1, p2, subsample_or_initxy1=4, device='cpu')
    b1, b2 = fast_nn.fast_reciprocal_NNs(p1, p2, subsample_or_initxy1=4, device='cpu', dist='l2')
    assert (a1 == b1).all() and (a2 == b2).all()


def test_extract_correspondences_nonsym_golden(golden):
    """``mast3r/fast_nn.py:191-223`` run by the live reference (oracle/gen_golden.py --fast-nn-extra)."""
    from oracle import fast_nn
    g = golden('fast_nn_extra.npz')
    d1, d2 = torch.from_numpy(g['d1']), torch.from_numpy(g['d2'])
    for tag, tol in (('tol0', 0), ('tol2', 2)):
        xy1, xy2, conf = fast_nn.extract_correspondences_nonsym(d1, d2, g['cA'], g['cB'], subsample=8, device='cpu',
                                                                pixel_tol=tol)
        assert (xy1.numpy() == g[f'{tag}/xy1']).all() and (xy2.numpy() == g[f'{tag}/xy2']).all()
        assert np.array_equal(conf.numpy(), g[f'{tag}/conf'])


def test_point_cloud_to_depth_oracle(golden):
    """Oracle restatement of ``point_cloud_to_depth`` against the live reference's outputs (row f4)."""
    g = golden('depth_splat.npz')
    for name in ('scene', 'sparse', 'none', 'halfpix'):
        w, h = (int(v) for v in g[f'{name}/wh'])
        out = functions.point_cloud_to_depth(T(g[f'{name}/pts']), T(g[f'{name}/K']), w, h)
        ref = g[f'{name}/depth']
        assert out.shape == ref.shape
        assert ((out.numpy() > 0) == (ref > 0)).all(), name          # same pixels hit: the rounding is bit-exact
        np.testing.assert_allclose(out.numpy(), ref, rtol=1e-6, atol=0, err_msg=name)


def _bf16_ulps(a, b):
    """Distance in bf16 steps between two bf16-representable fp32 arrays of positive numbers."""
    ia = (a.astype(np.float32).view(np.uint32) >> 16).astype(np.int64)
    ib = (b.astype(np.float32).view(np.uint32) >> 16).astype(np.int64)
    return np.abs(ia - ib)


def test_vggt_block_attention_oracle(golden):
    """``oracle.teacher.vggt_block_attention`` against the live ``vggt.layers.attention.Attention`` (row f2, VGGT)."""
    from oracle import teacher
    g = golden('vggt_attn.npz')
    for name in ('small', 'ragged_t3', 'blocks'):
        B, heads, n, L = (int(v) for v in g[f'{name}/meta'])
        temp, scale = float(g[f'{name}/temperature']), float(g[f'{name}/scale'])
        maps = []
        for blk in range(L):
            q, k = T(g[f'{name}/q{blk}']), T(g[f'{name}/k{blk}'])
            a32 = teacher.vggt_block_attention(q, k, scale, temp)
            np.testing.assert_allclose(a32.numpy(), g[f'{name}/attn_fp32_{blk}'], rtol=2e-6, atol=1e-9)
            # bf16 tensors: the score roundings are the reference's; its softmax output is rounded to bf16 as well
            a16 = teacher.vggt_block_attention(q.bfloat16(), k.bfloat16(), scale, temp)
            assert a16.dtype == torch.float32
            ulps = _bf16_ulps(a16.bfloat16().float().numpy(), g[f'{name}/attn_bf16_{blk}'])
            assert ulps.max() <= 1 and (ulps > 0).mean() < 0.01, (name, ulps.max(), (ulps > 0).mean())
            maps.append(a32)
        c1, c2 = teacher.vggt_cost_volumes(maps)
        np.testing.assert_allclose(c1.numpy(), g[f'{name}/cost_1'], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(c2.numpy(), g[f'{name}/cost_2'], rtol=2e-6, atol=1e-9)


def test_teacher_volume_oracle(golden):
    """``oracle.teacher.teacher_volume`` against ``tgt_attn_map`` of the live MASt3R teacher class (a small
    random-weight model run by ``oracle/gen_golden.py --teacher-volume``; row f2)."""
    from oracle import teacher
    g = golden('teacher_volume.npz')
    for name in ('recip', 'recip_t1', 'plain'):
        tgt, src = list(T(g[f'{name}/tgt'])), list(T(g[f'{name}/src']))
        out = teacher.teacher_volume(tgt, src, float(g[f'{name}/temperature']), bool(g[f'{name}/reciprocity']))
        np.testing.assert_allclose(out.numpy(), g[f'{name}/tgt_attn_map'], rtol=1e-6, atol=1e-9, err_msg=name)


LIVE_TAGS = ('mast3r0', 'mast3r1', 'vggt0', 'vggt1')


def live_case(g, tag):
    """Inputs of one ``live_bodies.npz`` case as the oracle / ops take them (shared with the GPU test)."""
    from oracle import synth
    ph, pw, C, K = (int(v) for v in g['meta'])
    H, W = ph * 14, pw * 14
    variant = tag[:-1]
    c = {k: T(g[f'{tag}/{k}']) for k in ('f1', 'f2', 't12', 't21', 'kp1', 'kp2', 'd1', 'd2', 'kf1', 'kf2', 'p3d1', 'p3d2',
                                         'dm1', 'dm2')}
    if g.has(f'{tag}/pixmask1'):      # co-visibility pixel masks -> patches by nearest sampling (finetune_timm_vggt.py:504-508)
        c['m1'] = T(g[f'{tag}/pixmask1'])[::14, ::14].reshape(-1)
        c['m2'] = T(g[f'{tag}/pixmask2'])[::14, ::14].reshape(-1)
    else:
        c['m1'] = functions.get_patch_mask_from_kp_tensor(c['kp1'][0], H, W, 14)
        c['m2'] = functions.get_patch_mask_from_kp_tensor(c['kp2'][0], H, W, 14)
    c['kd1'] = functions.extract_kp_depth(c['dm1'], c['kp1'])
    c['kd2'] = functions.extract_kp_depth(c['dm2'], c['kp2'])
    c['head_params'] = synth.head_params(int(g[f'{tag}/head_case']), C)
    c['variant'], c['C'] = variant, C
    return c


@pytest.mark.parametrize('tag', LIVE_TAGS)
def test_bodies_against_live_lightning_methods(golden, tag):
    """``oracle/bodies.py`` against losses and gradients of the reference's own ``calculate_cost_loss`` /
    ``calculate_matching_loss`` / ``calculate_depth_loss`` methods (``oracle/gen_live_bodies.py``)."""
    from oracle import bodies, synth
    g = golden('live_bodies.npz')
    c = live_case(g, tag)
    f1, f2 = c['f1'].clone().requires_grad_(True), c['f2'].clone().requires_grad_(True)
    kl = bodies.cost_volume_kl(f1, f2, c['t12'], c['t21'], c['m1'], c['m2'], c['variant'])
    kl.backward()
    assert rel_err(kl, g[f'{tag}/kl']) < 1e-5
    assert_grad_close(f1.grad, T(g[f'{tag}/grad_f1']), cos_min=0.99999, name='f1', norm_rtol=1e-3)
    assert_grad_close(f2.grad, T(g[f'{tag}/grad_f2']), cos_min=0.99999, name='f2', norm_rtol=1e-3)

    d1, d2 = c['d1'].clone().requires_grad_(True), c['d2'].clone().requires_grad_(True)
    ap = bodies.smooth_ap(d1[0], d2[0], c['p3d1'], c['p3d2'], c['variant'])
    ap.backward()
    assert rel_err(ap, g[f'{tag}/ap']) < 1e-5
    assert_grad_close(d1.grad, T(g[f'{tag}/grad_d1']), cos_min=0.99999, name='d1', norm_rtol=1e-3)
    assert_grad_close(d2.grad, T(g[f'{tag}/grad_d2']), cos_min=0.99999, name='d2', norm_rtol=1e-3)

    head = losses.DepthHead(c['C'])
    synth.load_head(head, c['head_params'])
    kf1, kf2 = c['kf1'].clone().requires_grad_(True), c['kf2'].clone().requires_grad_(True)
    l1, rank = bodies.depth_losses(head, kf1, kf2, c['kd1'], c['kd2'])
    (l1 + rank).backward()
    assert rel_err(l1, g[f'{tag}/l1']) < 1e-5 and rel_err(rank, g[f'{tag}/rank']) < 1e-5
    assert_grad_close(kf1.grad, T(g[f'{tag}/grad_kf1']), cos_min=0.99999, name='kf1', norm_rtol=1e-3)
    assert_grad_close(kf2.grad, T(g[f'{tag}/grad_kf2']), cos_min=0.99999, name='kf2', norm_rtol=1e-3)
    fl = head.fusion_layer
    packed = torch.cat([q.grad.reshape(-1) for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight,
                                                     fl[3].bias)])
    assert_grad_close(packed, T(g[f'{tag}/grad_head']), cos_min=0.99999, name='head', norm_rtol=1e-3)


@pytest.mark.parametrize('tag', ['me0', 'me1'])
def test_me_smooth_ap_against_live_training_step(golden, tag):
    """'me' variant (3-D positives, several per row) against the live ``FinetuneTIMM.training_step``."""
    g = golden('live_bodies.npz')
    d1, d2 = T(g[f'{tag}/d1']).clone().requires_grad_(True), T(g[f'{tag}/d2']).clone().requires_grad_(True)
    ap = bodies.smooth_ap(d1[0], d2[0], T(g[f'{tag}/p3d1']), T(g[f'{tag}/p3d2']), 'me')
    ap.backward()
    assert rel_err(ap, g[f'{tag}/ap']) < 1e-5
    assert_grad_close(d1.grad, T(g[f'{tag}/grad_d1']), cos_min=0.99999, name='d1', norm_rtol=1e-3)
    assert_grad_close(d2.grad, T(g[f'{tag}/grad_d2']), cos_min=0.99999, name='d2', norm_rtol=1e-3)


@pytest.mark.parametrize('tag', ['sample0', 'sample1'])
def test_sampling_glue_against_live_feature_getters(golden, tag):
    """``bodies.sample_tokens`` against the live ``get_intermediate_feature`` (4 layers sampled separately, then averaged)
    and ``get_feature`` (final features, L2-normalised) of ``FinetuneMASt3RTIMM`` run on a stand-in ViT."""
    g = golden('live_bodies.npz')
    gh, gw = (int(v) for v in g[f'{tag}/grid'])
    layers = T(g[f'{tag}/layers']).clone().requires_grad_(True)       # (4, N, C)
    final = T(g[f'{tag}/final']).clone().requires_grad_(True)
    kp = T(g[f'{tag}/kp'])
    feat = torch.stack([bodies.sample_tokens(layers[l][None], gh, gw, kp) for l in range(4)]).mean(dim=0)
    desc = bodies.sample_tokens(final[None], gh, gw, kp, normalize=True)
    np.testing.assert_allclose(feat.detach().numpy(), g[f'{tag}/feat'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(desc.detach().numpy(), g[f'{tag}/desc'], rtol=1e-5, atol=1e-6)
    ((feat * T(g[f'{tag}/w_feat'])).sum() + (desc * T(g[f'{tag}/w_desc'])).sum()).backward()
    assert_grad_close(layers.grad, T(g[f'{tag}/grad_layers']), cos_min=0.99999, name='layers', norm_rtol=1e-3)
    assert_grad_close(final.grad, T(g[f'{tag}/grad_final']), cos_min=0.99999, name='final', norm_rtol=1e-3)


def test_semantic_argmax_oracle_against_live_semantic_transfer(golden):
    """``oracle.evaluate.semantic_argmax`` against ``nn_idx`` recorded inside the live ``semantic_transfer``
    (``src/evaluate_timm.py:461-588``, run by ``oracle/gen_live_bodies.py --eval``; row f3)."""
    from oracle import evaluate
    g = golden('eval_argmax.npz')
    img, patch, stride, ph, C, K = (int(v) for v in g['meta'])
    d1 = T(g['tokens1']).reshape(1, ph, ph, C).permute(0, 3, 1, 2)
    d2 = T(g['tokens2']).reshape(1, ph, ph, C).permute(0, 3, 1, 2)
    # the live call passes h = w = img_size and leaves patch_size / stride at their defaults (:542)
    kd = functions.interpolate_features(d1, T(g['kps1'])[None, :, :2], h=img, w=img, normalize=True)
    idx, sim = evaluate.semantic_argmax(kd, d2, img, patch, stride)
    assert (idx.numpy() == g['nn_idx']).all()
    np.testing.assert_allclose(sim.max(dim=1).values.numpy(), g['best'], rtol=1e-6)


@pytest.mark.parametrize('tag', ['kpmatch0', 'kpmatch1'])
def test_keypoint_selection_oracle_against_live_method(golden, tag):
    from oracle import keypoints
    g = golden('live_bodies.npz')
    kp1, kp2 = keypoints.filter_and_match_keypoints(T(g[f'{tag}/desc1_x8']).float() / 8, T(g[f'{tag}/desc2_x8']).float() / 8,
                                                    T(g[f'{tag}/conf1']), T(g[f'{tag}/conf2']),
                                                    float(g[f'{tag}/min_conf_thr']))
    assert torch.equal(kp1, T(g[f'{tag}/kp1'])) and torch.equal(kp2, T(g[f'{tag}/kp2']))


@pytest.mark.parametrize('which', ['step', 'vstep'])
def test_oracle_whole_step_against_live_training_step(golden, which):
    """The oracle pieces chained like the reference's ``training_step`` reproduce the live step (both modules): total,
    partial losses and the gradients w.r.t. every ViT token and the depth head."""
    from oracle import keypoints
    g = golden('live_bodies.npz')
    ph, pw, C, _ = (int(v) for v in g['meta'])
    H, W = ph * 14, pw * 14
    variant = 'mast3r' if which == 'step' else 'vggt'
    layers = T(g['step/layers']).clone().requires_grad_(True)          # (view, block, N, C)
    final = T(g['step/final']).clone().requires_grad_(True)
    (pts1, z1), (pts2, z2) = synth.analytic_scene(H, W, 0), synth.analytic_scene(H, W, 1)
    head = losses.DepthHead(C)
    synth.load_head(head, synth.head_params(4400 if which == 'step' else 4401, C))
    if which == 'step':
        w_ap, w_depth, w_intra, w_kl = (float(v) for v in g['step/weights'])
        kp1, kp2 = keypoints.filter_and_match_keypoints(T(g['step/desc1_x8']).float() / 8, T(g['step/desc2_x8']).float() / 8,
                                                        T(g['step/conf1']), T(g['step/conf2']), float(g['step/min_conf_thr']))
        m1 = functions.get_patch_mask_from_kp_tensor(kp1[0], H, W, 14)
        m2 = functions.get_patch_mask_from_kp_tensor(kp2[0], H, W, 14)
        f1, f2 = layers[0].mean(dim=0), layers[1].mean(dim=0)           # mean of blocks 4..7
    else:
        w_ap = w_depth = w_intra = w_kl = 1.0
        kp1, kp2 = T(g['vstep/kp1']), T(g['vstep/kp2'])
        m1 = T(g['vstep/pixmask1'])[::14, ::14].reshape(-1)
        m2 = T(g['vstep/pixmask2'])[::14, ::14].reshape(-1)
        f1, f2 = layers[0, 3], layers[1, 3]                             # block 7 only
    kl = bodies.cost_volume_kl(f1, f2, T(g[f'{which}/cost1']), T(g[f'{which}/cost2']), m1, m2, variant)
    kf1 = torch.stack([bodies.sample_tokens(layers[0, l][None], ph, pw, kp1) for l in range(4)]).mean(dim=0)
    kf2 = torch.stack([bodies.sample_tokens(layers[1, l][None], ph, pw, kp2) for l in range(4)]).mean(dim=0)
    l1, rank = bodies.depth_losses(head, kf1, kf2, functions.extract_kp_depth(z1, kp1), functions.extract_kp_depth(z2, kp2))
    d1 = bodies.sample_tokens(final[0:1], ph, pw, kp1, normalize=True)[0]
    d2 = bodies.sample_tokens(final[1:2], ph, pw, kp2, normalize=True)[0]
    p1 = pts1[kp1[0, :, 1].long(), kp1[0, :, 0].long()]
    p2 = pts2[kp2[0, :, 1].long(), kp2[0, :, 0].long()]
    ap = bodies.smooth_ap(d1, d2, p1, p2, variant)
    loss = w_ap * ap + w_depth * l1 + w_intra * rank + w_kl * kl
    loss.backward()
    for got, ref in zip((ap, l1, rank, kl), g[f'{which}/parts']):
        assert rel_err(got, ref) < 1e-5
    assert rel_err(loss, g[f'{which}/loss']) < 1e-5
    assert_grad_close(layers.grad, T(g[f'{which}/grad_layers']), cos_min=0.99999, name='block tokens', norm_rtol=1e-3)
    assert_grad_close(final.grad, T(g[f'{which}/grad_final']), cos_min=0.99999, name='final tokens', norm_rtol=1e-3)
    fl = head.fusion_layer
    packed = torch.cat([q.grad.reshape(-1) for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight,
                                                     fl[3].bias)])
    assert_grad_close(packed, T(g[f'{which}/grad_head']), cos_min=0.99999, name='head', norm_rtol=1e-3)
