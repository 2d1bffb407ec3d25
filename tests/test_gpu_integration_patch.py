"""The reference-side patch of INTEGRATION.md section 3, executed: the three patched LightningModule methods exactly as
the document shows them, on a stand-in module whose feature getters return the golden features, against the losses and
gradients of the reference's own unpatched methods (``tests/golden/live_bodies.npz``)."""
import types

import pytest
import torch

from helpers import assert_grad_close, rel_err
from oracle import losses as olosses
from oracle import synth
from test_oracle_golden import LIVE_TAGS, live_case

pytestmark = pytest.mark.gpu
T = torch.as_tensor


# ---------------------------------------------------------------- INTEGRATION.md section 3, verbatim
def calculate_cost_loss(self, rgb_1_resized, rgb_2_resized, kp_1, kp_2, mast3r_cost_1, mast3r_cost_2, batch_idx):
    from gd3 import ops
    from gd3.compat.functions import get_patch_mask_from_kp_tensor
    feat_cost_1 = self.get_feature_cost(rgb_1_resized, normalize=False, resize=False)
    feat_cost_2 = self.get_feature_cost(rgb_2_resized, normalize=False, resize=False)
    B, _, H, W = rgb_1_resized.shape
    n = (H // self.patch_size) * (W // self.patch_size)
    mask_1 = get_patch_mask_from_kp_tensor(kp_1[0], H, W, self.patch_size)
    mask_2 = get_patch_mask_from_kp_tensor(kp_2[0], H, W, self.patch_size)
    return ops.cost_volume_kl(feat_cost_1.view(1, n, -1), feat_cost_2.view(1, n, -1),
                              mast3r_cost_1[None], mast3r_cost_2[None], mask_1, mask_2, variant=self.variant)[0]


def calculate_matching_loss(self, rgb_1_resized, rgb_2_resized, kp_1, kp_2, pts3d_1_map, pts3d_2_map):
    from gd3 import ops
    desc_1 = self.get_feature(rgb_1_resized, kp_1, normalize=True)
    desc_2 = self.get_feature(rgb_2_resized, kp_2, normalize=True)
    pts3d_1 = pts3d_1_map[kp_1[..., 1].long(), kp_1[..., 0].long()]
    pts3d_2 = pts3d_2_map[kp_2[..., 1].long(), kp_2[..., 0].long()]
    return ops.smooth_ap(desc_1, desc_2, pts3d_1, pts3d_2, variant=self.variant, thr_neg=self.thres3d_neg)[0]


def calculate_depth_loss(self, depth_pred_1, depth_pred_2, rgb_1_resized, rgb_2_resized, kp_1, kp_2, indices=[4, 5, 6, 7]):
    from gd3 import ops
    from gd3.compat.functions import extract_kp_depth
    kp_feat_1 = self.get_intermediate_feature(rgb_1_resized, n=indices, pts=kp_1, reshape=True, normalize=True)
    kp_feat_2 = self.get_intermediate_feature(rgb_2_resized, n=indices, pts=kp_2, reshape=True, normalize=True)
    kp_depth_1 = extract_kp_depth(depth_pred_1, kp_1)
    kp_depth_2 = extract_kp_depth(depth_pred_2, kp_2)
    feats = torch.cat([kp_feat_1, kp_feat_2])            # sets (view 1, view 2)
    depths = torch.cat([kp_depth_1, kp_depth_2])
    one = feats.new_ones(1)
    total, loss_rank, loss_l1 = ops.depth_head_loss(self.depth_diff_head, feats, depths, depth_threshold=0.05,
                                                    w_rank=one.expand(2) * 0.5, w_l1=one)
    return total, loss_l1[0], 0.5 * (loss_rank[0] + loss_rank[1])
# ----------------------------------------------------------------


@pytest.mark.parametrize('tag', [t for t in LIVE_TAGS if t != 'vggt1'])      # vggt1 uses pixel masks, not keypoints
def test_patched_methods_match_the_live_methods(golden, tag):
    g = golden('live_bodies.npz')
    c = live_case(g, tag)
    ph, pw, C, K = (int(v) for v in g['meta'])
    H, W = ph * 14, pw * 14
    dev = 'cuda'
    rgb1, rgb2 = torch.zeros(1, 3, H, W, device=dev), torch.zeros(1, 3, H, W, device=dev)
    f = {id(rgb1): c['f1'].to(dev).requires_grad_(True), id(rgb2): c['f2'].to(dev).requires_grad_(True)}
    d = {id(rgb1): c['d1'].to(dev).requires_grad_(True), id(rgb2): c['d2'].to(dev).requires_grad_(True)}
    kf = {id(rgb1): c['kf1'].to(dev).requires_grad_(True), id(rgb2): c['kf2'].to(dev).requires_grad_(True)}
    head = olosses.DepthHead(C)
    synth.load_head(head, c['head_params'])
    me = types.SimpleNamespace(
        variant=c['variant'], patch_size=14, thres3d_neg=0.1, depth_diff_head=head.to(dev),
        # what the live getter returns (src/finetune_timm_mast3r.py:322-337): a channel-major (1, ph, pw, C) view
        get_feature_cost=lambda rgb, normalize=False, resize=False:
            f[id(rgb)].reshape(1, ph, pw, C).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1),
        get_feature=lambda rgb, kp, normalize=True: d[id(rgb)],
        get_intermediate_feature=lambda rgb, n=None, pts=None, reshape=True, normalize=True: kf[id(rgb)])
    kp1, kp2 = c['kp1'].to(dev), c['kp2'].to(dev)

    kl = calculate_cost_loss(me, rgb1, rgb2, kp1, kp2, c['t12'].to(dev), c['t21'].to(dev), 0)
    kl.backward()
    assert rel_err(kl.detach().cpu(), g[f'{tag}/kl']) < 1e-3
    assert_grad_close(f[id(rgb1)].grad.cpu(), T(g[f'{tag}/grad_f1']), name='f1', norm_rtol=3e-2)
    assert_grad_close(f[id(rgb2)].grad.cpu(), T(g[f'{tag}/grad_f2']), name='f2', norm_rtol=3e-2)

    pm1, pm2 = torch.zeros(H, W, 3, device=dev), torch.zeros(H, W, 3, device=dev)
    pm1[kp1[0, :, 1].long(), kp1[0, :, 0].long()] = c['p3d1'].to(dev)
    pm2[kp2[0, :, 1].long(), kp2[0, :, 0].long()] = c['p3d2'].to(dev)
    ap = calculate_matching_loss(me, rgb1, rgb2, kp1, kp2, pm1, pm2)
    ap.backward()
    assert rel_err(ap.detach().cpu(), g[f'{tag}/ap']) < 1e-3
    assert_grad_close(d[id(rgb1)].grad.cpu(), T(g[f'{tag}/grad_d1']), name='d1', norm_rtol=3e-2)

    total, l1, rank = calculate_depth_loss(me, c['dm1'].to(dev), c['dm2'].to(dev), rgb1, rgb2, kp1, kp2)
    total.backward()
    assert rel_err(l1.cpu(), g[f'{tag}/l1']) < 1e-3 and rel_err(rank.cpu(), g[f'{tag}/rank']) < 1e-3
    assert_grad_close(kf[id(rgb1)].grad.cpu(), T(g[f'{tag}/grad_kf1']), name='kf1', norm_rtol=3e-2)
    assert_grad_close(kf[id(rgb2)].grad.cpu(), T(g[f'{tag}/grad_kf2']), name='kf2', norm_rtol=3e-2)
