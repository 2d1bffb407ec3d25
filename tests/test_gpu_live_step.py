"""One whole training step through the CUDA path against the reference's own ``FinetuneMASt3RTIMM.training_step``
run live (``oracle/gen_live_bodies.py``; only the teacher and the ViT were stand-ins returning the fixed maps / tokens
stored in ``tests/golden/live_bodies.npz``): teacher-side keypoint matching and filtering, keypoint depths, the three
feature samplings, KL + Smooth-AP + depth losses, the weighted sum, and the gradients w.r.t. every ViT token and the
depth head."""
import pytest
import torch

from helpers import assert_grad_close, rel_err
from oracle import losses as olosses
from oracle import synth

pytestmark = pytest.mark.gpu
T = torch.as_tensor


def test_whole_step_vs_live_training_step(golden):
    from gd3 import ops
    from gd3.compat import functions as cfn
    from gd3.compat import keypoints
    g = golden('live_bodies.npz')
    ph, pw, C, _ = (int(v) for v in g['meta'])
    H, W = ph * 14, pw * 14
    w_ap, w_depth, w_intra, w_kl = (float(v) for v in g['step/weights'])
    layers = T(g['step/layers']).cuda().requires_grad_(True)          # (view, 4 blocks, N, C) patch tokens
    final = T(g['step/final']).cuda().requires_grad_(True)            # (view, N, C)
    head = olosses.DepthHead(C)
    synth.load_head(head, synth.head_params(4400, C))
    head = head.cuda()

    # 2. teacher-side keypoints (src/finetune_timm_mast3r.py:392-469)
    teacher = dict(desc_1=T(g['step/desc1_x8']).float() / 8, desc_2=T(g['step/desc2_x8']).float() / 8,
                   conf_1=T(g['step/conf1']), conf_2=T(g['step/conf2']))
    kp1, kp2, w, h = keypoints.filter_and_match_keypoints(teacher, float(g['step/min_conf_thr']))
    assert (w, h) == (W, H) and kp1.shape[1] > 8
    (pts1, z1), (pts2, z2) = (tuple(t.cuda() for t in synth.analytic_scene(H, W, v)) for v in (0, 1))

    # 3. depth losses (:472-501): features = mean of the 4 sampled blocks, depths = 3 x 3 windows of the depth maps
    kf1 = ops.sample_tokens(layers[0][:, None], (ph, pw), kp1)
    kf2 = ops.sample_tokens(layers[1][:, None], (ph, pw), kp2)
    kd = torch.cat([cfn.extract_kp_depth(z1, kp1), cfn.extract_kp_depth(z2, kp2)])
    one = kf1.new_ones(1)
    depth_total, rank, l1 = ops.depth_head_loss(head, torch.cat([kf1, kf2]), kd, depth_threshold=0.05,
                                                w_rank=one.expand(2) * 0.5 * w_intra, w_l1=one * w_depth)

    # 4. cost-volume KL (:504-540): features = mean of the 4 blocks, masks = patches holding a keypoint
    m1 = cfn.get_patch_mask_from_kp_tensor(kp1[0], H, W, 14)
    m2 = cfn.get_patch_mask_from_kp_tensor(kp2[0], H, W, 14)
    kl = ops.cost_volume_kl(layers[0].mean(dim=0)[None], layers[1].mean(dim=0)[None], T(g['step/cost1']).cuda()[None],
                            T(g['step/cost2']).cuda()[None], m1, m2, variant='mast3r')[0]

    # 5. Smooth-AP (:543-589): normalised samples of the final features, 3-D points read at the keypoints
    d1 = ops.sample_tokens(final[0:1], (ph, pw), kp1, normalize=True)
    d2 = ops.sample_tokens(final[1:2], (ph, pw), kp2, normalize=True)
    p1 = pts1[kp1[..., 1].long(), kp1[..., 0].long()]
    p2 = pts2[kp2[..., 1].long(), kp2[..., 0].long()]
    ap = ops.smooth_ap(d1, d2, p1, p2, variant='mast3r')[0]

    # 6. total (:650-653)
    loss = w_ap * ap + depth_total + w_kl * kl
    loss.backward()

    ap_ref, depth_ref, intra_ref, kl_ref = (float(v) for v in g['step/parts'])
    assert rel_err(ap.detach().cpu(), ap_ref) < 1e-3
    assert rel_err(l1[0].cpu(), depth_ref) < 1e-3
    assert rel_err((0.5 * (rank[0] + rank[1])).cpu(), intra_ref) < 1e-3
    assert rel_err(kl.detach().cpu(), kl_ref) < 1e-3
    assert rel_err(loss.detach().cpu(), g['step/loss']) < 1e-3
    assert_grad_close(layers.grad.cpu(), T(g['step/grad_layers']), name='block tokens', norm_rtol=3e-2)
    assert_grad_close(final.grad.cpu(), T(g['step/grad_final']), name='final tokens', norm_rtol=3e-2)
    fl = head.fusion_layer
    packed = torch.cat([q.grad.reshape(-1) for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight,
                                                     fl[3].bias)]).cpu()
    assert_grad_close(packed, T(g['step/grad_head']), name='head', norm_rtol=3e-2)


def test_whole_vggt_step_vs_live_training_step(golden):
    """Same for ``FinetuneVGGTTIMM.training_step`` (src/finetune_timm_vggt.py:577-639): KL on block 7 only with
    co-visibility pixel masks (variant 'vggt'), the vggt Smooth-AP variant, keypoints handed over by the teacher."""
    from gd3 import ops
    from gd3.compat import functions as cfn
    g = golden('live_bodies.npz')
    ph, pw, C, _ = (int(v) for v in g['meta'])
    H, W = ph * 14, pw * 14
    layers = T(g['step/layers']).cuda().requires_grad_(True)          # same token values as the MASt3R step
    final = T(g['step/final']).cuda().requires_grad_(True)
    head = olosses.DepthHead(C)
    synth.load_head(head, synth.head_params(4401, C))
    head = head.cuda()
    kp1, kp2 = T(g['vstep/kp1']).cuda(), T(g['vstep/kp2']).cuda()
    (pts1, z1), (pts2, z2) = (tuple(t.cuda() for t in synth.analytic_scene(H, W, v)) for v in (0, 1))

    kf1 = ops.sample_tokens(layers[0][:, None], (ph, pw), kp1)
    kf2 = ops.sample_tokens(layers[1][:, None], (ph, pw), kp2)
    kd = torch.cat([cfn.extract_kp_depth(z1, kp1), cfn.extract_kp_depth(z2, kp2)])
    one = kf1.new_ones(1)
    depth_total, rank, l1 = ops.depth_head_loss(head, torch.cat([kf1, kf2]), kd, depth_threshold=0.05,
                                                w_rank=one.expand(2) * 0.5, w_l1=one)

    # pixel masks -> patches by nearest sampling (:504-508); student features = block 7 (:342)
    m1 = T(g['vstep/pixmask1'])[::14, ::14].reshape(-1).cuda()
    m2 = T(g['vstep/pixmask2'])[::14, ::14].reshape(-1).cuda()
    kl = ops.cost_volume_kl(layers[0, 3][None], layers[1, 3][None], T(g['vstep/cost1']).cuda()[None],
                            T(g['vstep/cost2']).cuda()[None], m1, m2, variant='vggt')[0]

    d1 = ops.sample_tokens(final[0:1], (ph, pw), kp1, normalize=True)
    d2 = ops.sample_tokens(final[1:2], (ph, pw), kp2, normalize=True)
    p1 = pts1[kp1[..., 1].long(), kp1[..., 0].long()]
    p2 = pts2[kp2[..., 1].long(), kp2[..., 0].long()]
    ap = ops.smooth_ap(d1, d2, p1, p2, variant='vggt')[0]

    loss = ap + depth_total + kl
    loss.backward()
    ap_ref, depth_ref, intra_ref, kl_ref = (float(v) for v in g['vstep/parts'])
    assert rel_err(ap.detach().cpu(), ap_ref) < 1e-3
    assert rel_err(l1[0].cpu(), depth_ref) < 1e-3
    assert rel_err((0.5 * (rank[0] + rank[1])).cpu(), intra_ref) < 1e-3
    assert rel_err(kl.detach().cpu(), kl_ref) < 1e-3
    assert rel_err(loss.detach().cpu(), g['vstep/loss']) < 1e-3
    assert_grad_close(layers.grad.cpu(), T(g['vstep/grad_layers']), name='block tokens', norm_rtol=3e-2)
    assert_grad_close(final.grad.cpu(), T(g['vstep/grad_final']), name='final tokens', norm_rtol=3e-2)
    fl = head.fusion_layer
    packed = torch.cat([q.grad.reshape(-1) for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight,
                                                     fl[3].bias)]).cpu()
    assert_grad_close(packed, T(g['vstep/grad_head']), name='head', norm_rtol=3e-2)
