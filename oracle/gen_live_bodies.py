"""Golden vectors from the LIVE loss methods of the reference's LightningModules -> ``tests/golden/live_bodies.npz``.

Run in the build container only (``python -m oracle.gen_live_bodies``); needs ``/root/reference``.

``src/finetune_timm_mast3r.py`` and ``src/finetune_timm_vggt.py`` import packages this image does not have (timm,
pytorch_lightning, hydra, visdom, matplotlib, albumentations, kornia ...).  None of them is touched by the three loss
methods, so a meta-path finder registers empty stand-in modules for exactly those absent packages (real modules always
win: the finder is appended), the modules are imported unmodified, and

    FinetuneMASt3RTIMM.calculate_cost_loss / calculate_matching_loss / calculate_depth_loss
    FinetuneVGGTTIMM.calculate_cost_loss   / calculate_matching_loss / calculate_depth_loss
    FinetuneTIMM.training_step             (src/finetune_timm_me.py: the ME baseline's Smooth-AP with 3-D positives)
    FinetuneMASt3RTIMM.get_intermediate_feature / get_feature   (keypoint sampling glue, stand-in ViT)
    FinetuneMASt3RTIMM.filter_and_match_keypoints               (teacher-side keypoints: reciprocal NN + filters)
    FinetuneMASt3RTIMM.training_step                            (the whole step; teacher and ViT are stand-ins)
    FinetuneVGGTTIMM.training_step                              (likewise; the teacher's tracker picks the keypoints)
    evaluate_timm.semantic_transfer                             (--eval -> tests/golden/eval_argmax.npz)

are called as plain functions on a stand-in ``self`` that supplies what they read from the module: the ViT feature
getters (returning the given synthetic features instead of running a backbone), ``depth_diff_head`` (the live
``DepthAwareFeatureFusion``), ``patch_size``, ``thres3d_neg``, ``device``.  Losses and autograd gradients of those live
bodies are stored next to their inputs; ``tests/test_oracle_golden.py`` pins ``oracle/bodies.py`` to them and the ``-m
gpu`` tests compare the CUDA ops with them.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('GD3_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

ABSENT = {'pytorch_lightning', 'timm', 'visdom', 'cv2', 'matplotlib', 'albumentations', 'hydra', 'omegaconf',
          'huggingface_hub', 'kornia', 'roma', 'trimesh', 'open3d', 'imageio', 'pycocotools', 'h5py', 'wandb',
          'tensorboard', 'lightning'}


def _stand_in_class(name):
    def call(self, *a, **k):          # usable as a decorator factory (hydra.main) and as a constructor
        return a[0] if len(a) == 1 and callable(a[0]) and not k else self
    return type(name, (object,), {'__init__': lambda self, *a, **k: None, '__call__': call,
                                  '__getattr__': lambda self, n: _stand_in_class(n)()})


class _StandInModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        cls = _stand_in_class(name)
        setattr(self, name, cls)
        return cls


class _AbsentPackages(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split('.')[0] in ABSENT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StandInModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_live_modules():
    sys.meta_path.append(_AbsentPackages())
    for p in (REF, os.path.join(REF, 'src'), os.path.join(REF, 'dust3r')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import finetune_timm_mast3r as ft_mast3r
    import finetune_timm_vggt as ft_vggt
    import finetune_timm_me as ft_me
    import utils.model as ref_model
    return ft_mast3r, ft_vggt, ft_me, ref_model


def _np(x):
    return x.detach().cpu().numpy()


class _Self:
    """What the loss methods read from the LightningModule."""
    device = 'cpu'
    patch_size = 14
    resize_patch_size = 14
    thres3d_neg = 0.1

    def __init__(self, by_image):
        self.by_image = by_image            # id(rgb tensor) -> dict of the features that image "produces"

    def get_feature_cost(self, rgb, normalize=False, resize=False):
        return self.by_image[id(rgb)]['cost']

    def get_feature(self, rgb, kp, normalize=True):
        return self.by_image[id(rgb)]['desc']

    def get_intermediate_feature(self, rgb, n=None, pts=None, reshape=True, return_class_token=False, normalize=True):
        return self.by_image[id(rgb)]['kp_feat']


def main():
    from oracle import synth
    ft_mast3r, ft_vggt, ft_me, ref_model = import_live_modules()
    out = {}
    ph, pw, C, K = 8, 10, 64, 40
    H, W, N = ph * 14, pw * 14, ph * pw
    for variant, cls in (('mast3r', ft_mast3r.FinetuneMASt3RTIMM), ('vggt', ft_vggt.FinetuneVGGTTIMM)):
        for case in range(2):
            tag = f'{variant}{case}'
            it = synth.pair_inputs(9, 10 * case + (0 if variant == 'mast3r' else 5), N, C, K, (ph, pw), variant)
            rgb1, rgb2 = torch.zeros(1, 3, H, W), torch.zeros(1, 3, H, W)
            kp1, kp2 = it.kp1[None], it.kp2[None]
            leaves = dict(f1=it.f1.clone().requires_grad_(True), f2=it.f2.clone().requires_grad_(True))
            # descriptors as get_feature(normalize=True) returns them: unit rows, (1, K, C)
            d1 = torch.nn.functional.normalize(it.g1[:K], dim=-1)[None].clone().requires_grad_(True)
            d2 = torch.nn.functional.normalize(it.g1[:K] + 0.25 * it.g2[:K], dim=-1)[None].clone().requires_grad_(True)
            kf1 = it.g1[K:2 * K][None].clone().requires_grad_(True)
            kf2 = it.g2[K:2 * K][None].clone().requires_grad_(True)
            me = _Self({id(rgb1): dict(cost=leaves['f1'].view(1, ph, pw, C), desc=d1, kp_feat=kf1),
                        id(rgb2): dict(cost=leaves['f2'].view(1, ph, pw, C), desc=d2, kp_feat=kf2)})
            torch.manual_seed(77)
            me.depth_diff_head = ref_model.DepthAwareFeatureFusion(C)
            synth.load_head(me.depth_diff_head, synth.head_params(4300 + case, C))

            # ---- cost-volume KL ----
            if variant == 'mast3r':
                kl = cls.calculate_cost_loss(me, rgb1, rgb2, kp1, kp2, it.t12, it.t21, 0)
            elif case == 0:     # keypoint masks
                kl = cls.calculate_cost_loss(me, rgb1, rgb2, it.t12[None], it.t21[None], kp_1=kp1, kp_2=kp2)
            else:               # co-visibility pixel masks, reduced to patches by the method itself
                pix1 = it.m1.view(ph, pw).repeat_interleave(14, 0).repeat_interleave(14, 1)
                pix2 = it.m2.view(ph, pw).repeat_interleave(14, 0).repeat_interleave(14, 1)
                out[f'{tag}/pixmask1'], out[f'{tag}/pixmask2'] = _np(pix1), _np(pix2)
                kl = cls.calculate_cost_loss(me, rgb1, rgb2, it.t12[None], it.t21[None], mask_1=pix1, mask_2=pix2)
            kl.backward()

            # ---- Smooth-AP ----
            g = torch.Generator().manual_seed(500 + case)
            pm1 = torch.rand(H, W, 3, generator=g)
            pm2 = torch.rand(H, W, 3, generator=g)
            yx1, yx2 = (kp1[0, :, 1].long(), kp1[0, :, 0].long()), (kp2[0, :, 1].long(), kp2[0, :, 0].long())
            pm1[yx1] = it.p1
            pm2[yx2] = it.p2
            ap = cls.calculate_matching_loss(me, rgb1, rgb2, kp1, kp2, pm1, pm2)
            ap.backward()

            # ---- depth losses ----
            dm1 = torch.rand(H, W, generator=g) * 4 + 0.5
            dm2 = torch.rand(H, W, generator=g) * 4 + 0.5
            if variant == 'mast3r':
                l1, rank = cls.calculate_depth_loss(me, dm1, dm2, rgb1, rgb2, kp1, kp2)
            else:
                l1, rank = cls.calculate_depth_loss(me, dict(depth_pred_1=dm1, depth_pred_2=dm2), rgb1, rgb2, kp1, kp2)
            (l1 + rank).backward()

            fl = me.depth_diff_head.fusion_layer
            head_grads = torch.cat([p.grad.reshape(-1) for p in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias,
                                                                 fl[3].weight, fl[3].bias)])
            out.update({f'{tag}/f1': _np(it.f1), f'{tag}/f2': _np(it.f2), f'{tag}/t12': _np(it.t12), f'{tag}/t21': _np(it.t21),
                        f'{tag}/kp1': _np(kp1), f'{tag}/kp2': _np(kp2), f'{tag}/d1': _np(d1), f'{tag}/d2': _np(d2),
                        f'{tag}/kf1': _np(kf1), f'{tag}/kf2': _np(kf2), f'{tag}/p3d1': _np(pm1[yx1]), f'{tag}/p3d2': _np(pm2[yx2]),
                        f'{tag}/dm1': _np(dm1), f'{tag}/dm2': _np(dm2), f'{tag}/head_case': np.array(4300 + case),
                        f'{tag}/kl': _np(kl), f'{tag}/ap': _np(ap), f'{tag}/l1': _np(l1), f'{tag}/rank': _np(rank),
                        f'{tag}/grad_f1': _np(leaves['f1'].grad), f'{tag}/grad_f2': _np(leaves['f2'].grad),
                        f'{tag}/grad_d1': _np(d1.grad), f'{tag}/grad_d2': _np(d2.grad),
                        f'{tag}/grad_kf1': _np(kf1.grad), f'{tag}/grad_kf2': _np(kf2.grad),
                        f'{tag}/grad_head': _np(head_grads)})
            print(tag, 'kl', float(kl), 'ap', float(ap), 'l1', float(l1), 'rank', float(rank))
    # ---- the ME baseline's loss is the body of its training_step (src/finetune_timm_me.py:191-220): positives are all
    #      (s, t) closer than 5 mm in 3-D, so points are near-duplicated to give rows with several positives ----
    for case in range(2):
        tag = f'me{case}'
        it = synth.pair_inputs(9, 40 + case, N, C, K, (ph, pw), 'mast3r')
        g = torch.Generator().manual_seed(600 + case)
        p1 = torch.rand(K, 3, generator=g)
        p1[K // 2:K // 2 + 6] = p1[:6] + 0.001 * torch.randn(6, 3, generator=g)
        p2 = p1 + 0.0015 * torch.randn(K, 3, generator=g)
        d1 = torch.nn.functional.normalize(it.g1[:K], dim=-1)[None].clone().requires_grad_(True)
        d2 = torch.nn.functional.normalize(it.g1[:K] + 0.25 * it.g2[:K], dim=-1)[None].clone().requires_grad_(True)
        rgb1, rgb2 = torch.zeros(1, 3, H, W), torch.zeros(1, 3, H, W)
        me = _Self({id(rgb1): dict(desc=d1), id(rgb2): dict(desc=d2)})
        me.thresh3d_pos = 5e-3
        me.log = lambda *a, **k: None
        batch = dict(rgb_1=rgb1, pts2d_1=it.kp1[None], pts3d_1=p1[None], rgb_2=rgb2, pts2d_2=it.kp2[None], pts3d_2=p2[None])
        loss = ft_me.FinetuneTIMM.training_step(me, batch, 0)
        loss.backward()
        n_pos = int((torch.cdist(p1[None], p2[None]) < 5e-3).sum())
        out.update({f'{tag}/d1': _np(d1), f'{tag}/d2': _np(d2), f'{tag}/p3d1': _np(p1), f'{tag}/p3d2': _np(p2),
                    f'{tag}/ap': _np(loss), f'{tag}/grad_d1': _np(d1.grad), f'{tag}/grad_d2': _np(d2.grad)})
        print(tag, 'ap', float(loss), 'positives', n_pos)
    # ---- keypoint feature getters (src/finetune_timm_mast3r.py:242-313): the live methods with a stand-in ViT that
    #      returns given token tensors.  get_intermediate_feature samples 4 layers separately and averages them,
    #      get_feature samples the final features and L2-normalises ----
    class _ViT:
        num_prefix_tokens = 1
        norm = torch.nn.Identity()

        def __init__(self, layers, final):
            self.layers, self.final = layers, final

        def _intermediate_layers(self, x, n):
            return [self.layers[i] for i in range(len(n))]

        def forward_features(self, x):
            return self.final

    for case, (gh, gw) in enumerate([(8, 10), (9, 7)]):
        tag = f'sample{case}'
        Hs, Ws = gh * 14, gw * 14
        g = torch.Generator().manual_seed(700 + case)
        layers = [torch.randn(1, 1 + gh * gw, C, generator=g).requires_grad_(True) for _ in range(4)]
        final = torch.randn(1, 1 + gh * gw, C, generator=g).requires_grad_(True)
        kp = synth.keypoints(710 + case, K, Ws, Hs)[None]
        me = _Self({})
        me.model = _ViT(layers, final)
        me.input_transform = lambda x: x
        me.refine_conv = torch.nn.Identity()
        me.target_res = max(Hs, Ws)
        me.downsample_factor = 14
        rgb = torch.zeros(1, 3, Hs, Ws)
        feat = ft_mast3r.FinetuneMASt3RTIMM.get_intermediate_feature(me, rgb, pts=kp, n=[4, 5, 6, 7], reshape=True,
                                                                     return_class_token=False, normalize=True)
        desc = ft_mast3r.FinetuneMASt3RTIMM.get_feature(me, rgb, kp, normalize=True)
        w_f = torch.randn(feat.shape, generator=g)
        w_d = torch.randn(desc.shape, generator=g)
        ((feat * w_f).sum() + (desc * w_d).sum()).backward()
        out.update({f'{tag}/layers': _np(torch.stack([t[0, 1:] for t in layers])), f'{tag}/final': _np(final[0, 1:]),
                    f'{tag}/kp': _np(kp), f'{tag}/grid': np.array([gh, gw]), f'{tag}/feat': _np(feat), f'{tag}/desc': _np(desc),
                    f'{tag}/w_feat': _np(w_f), f'{tag}/w_desc': _np(w_d),
                    f'{tag}/grad_layers': _np(torch.stack([t.grad[0, 1:] for t in layers])),
                    f'{tag}/grad_final': _np(final.grad[0, 1:])})
        print(tag, tuple(feat.shape), tuple(desc.shape))
    # ---- teacher-side keypoints (src/finetune_timm_mast3r.py:392-469): live fast_reciprocal_NNs + border and confidence
    #      filters on exactly-representable descriptor maps ----
    for case, (mh, mw, thr) in enumerate([(64, 96, 10.0), (48, 80, 35.0)]):
        tag = f'kpmatch{case}'
        dmap1, dmap2 = synth.nn_desc_maps(800 + case, mh, mw, noise=0.5)
        dmap2 = torch.roll(dmap2, shifts=(3, -5), dims=(0, 1))          # view 2 is a shifted, noisier copy of view 1
        dmap1, dmap2 = torch.round(dmap1 * 16) / 8, torch.round(dmap2 * 16) / 8
        g = torch.Generator().manual_seed(810 + case)
        conf1, conf2 = torch.rand(mh, mw, generator=g) + 1.0, torch.rand(mh, mw, generator=g) + 1.0
        me = _Self({})
        me.min_conf_thr = thr
        feats = dict(view_1=dict(true_shape=torch.tensor([[mh, mw]])), view_2=dict(true_shape=torch.tensor([[mh, mw]])),
                     desc_1=dmap1, desc_2=dmap2, conf_1=conf1, conf_2=conf2)
        rgb = torch.zeros(1, 3, mh, mw)
        kp_1, kp_2, _, _, w_, h_ = ft_mast3r.FinetuneMASt3RTIMM.filter_and_match_keypoints(me, feats, rgb, rgb)
        out.update({f'{tag}/desc1_x8': _np(dmap1 * 8).astype(np.int8), f'{tag}/desc2_x8': _np(dmap2 * 8).astype(np.int8), f'{tag}/conf1': _np(conf1), f'{tag}/conf2': _np(conf2),
                    f'{tag}/min_conf_thr': np.array(thr, dtype=np.float32), f'{tag}/kp1': _np(kp_1), f'{tag}/kp2': _np(kp_2),
                    f'{tag}/wh': np.array([w_, h_])})
        print(tag, tuple(kp_1.shape))
    # ---- one whole training_step of FinetuneMASt3RTIMM (src/finetune_timm_mast3r.py:592-676), live: keypoint
    #      matching and filtering, the three feature getters, the three losses and their weighted sum.  Stand-ins: the
    #      teacher (extract_mast3r_features returns fixed maps) and the ViT (returns fixed token tensors per image) ----
    class _PairViT(_ViT):
        def __init__(self, per_image):
            self.per_image = per_image                      # 0 / 1 -> (layers, final)

        def _which(self, x):
            return int(float(x.mean()) > 0.5)               # image 1 is all zeros, image 2 all ones

        def _intermediate_layers(self, x, n):
            return [self.per_image[self._which(x)][0][i - 4] for i in n]      # blocks 4..7 are held

        def forward_features(self, x):
            return self.per_image[self._which(x)][1]

    mh, mw = ph * 14, pw * 14
    g = torch.Generator().manual_seed(900)
    # smooth token maps (a common component + per-token detail), view 2 = view 1 + noise: sampled descriptors of matched
    # keypoints correlate, so the Smooth-AP sigmoids are not all saturated
    def token_map(base=None):
        t = torch.randn(1, 1, C, generator=g) + 0.35 * torch.randn(1, 1 + N, C, generator=g) if base is None \
            else base.detach() + 0.08 * torch.randn(1, 1 + N, C, generator=g)
        return t.requires_grad_(True)
    first = ([token_map() for _ in range(4)], token_map())
    tokens = [first, ([token_map(t) for t in first[0]], token_map(first[1]))]
    dmap1, dmap2 = synth.nn_desc_maps(901, mh, mw, noise=0.5)
    dmap1, dmap2 = torch.round(dmap1 * 16) / 8, torch.roll(torch.round(dmap2 * 16) / 8, shifts=(2, -3), dims=(0, 1))
    conf1, conf2 = torch.rand(mh, mw, generator=g) + 1.0, torch.rand(mh, mw, generator=g) + 1.0
    (pts1, z1), (pts2, z2) = synth.analytic_scene(mh, mw, 0), synth.analytic_scene(mh, mw, 1)
    cost1, cost2 = synth.teacher_volume(902, N, 'mast3r'), synth.teacher_volume(903, N, 'mast3r')
    teacher_out = dict(view_1=dict(true_shape=torch.tensor([[mh, mw]])), view_2=dict(true_shape=torch.tensor([[mh, mw]])),
                       desc_1=dmap1, desc_2=dmap2, conf_1=conf1, conf_2=conf2, pts3d_1=pts1, pts3d_2_from_1=pts2,
                       pts3d_2=pts2, cost_1=cost1, cost_2=cost2)
    cls = ft_mast3r.FinetuneMASt3RTIMM
    me = _Self({})
    me.model = _PairViT(tokens)
    me.input_transform = lambda x: x
    me.refine_conv = torch.nn.Identity()
    me.target_res, me.downsample_factor, me.min_conf_thr = max(mh, mw), 14, 10.0
    me.ap_loss_weight, me.depth_loss_weight, me.intra_depth_loss_weight, me.kl_loss_weight = 1.0, 0.5, 1.0, 2.0
    torch.manual_seed(78)
    me.depth_diff_head = ref_model.DepthAwareFeatureFusion(C)
    synth.load_head(me.depth_diff_head, synth.head_params(4400, C))
    me.log = lambda *a, **k: None
    me.extract_mast3r_features = lambda a, b: teacher_out
    for name in ('filter_and_match_keypoints', 'calculate_depth_loss', 'calculate_cost_loss', 'calculate_matching_loss',
                 'get_intermediate_feature', 'get_feature', 'get_feature_cost'):
        setattr(me, name, types.MethodType(getattr(cls, name), me))
    batch = dict(rgb_1=torch.zeros(1, 3, mh, mw), rgb_2=torch.ones(1, 3, mh, mw), rgb_mast3r_1=None, rgb_mast3r_2=None,
                 intrinsic=torch.eye(3)[None], depth_1=z1[None], depth_2=z2[None])
    loss = cls.training_step(me, batch, 0)
    loss.backward()
    fl = me.depth_diff_head.fusion_layer
    out.update({'step/layers': _np(torch.stack([torch.stack([t[0, 1:] for t in tokens[v][0]]) for v in range(2)])),
                'step/final': _np(torch.stack([tokens[v][1][0, 1:] for v in range(2)])),
                'step/desc1_x8': _np(dmap1 * 8).astype(np.int8), 'step/desc2_x8': _np(dmap2 * 8).astype(np.int8),
                'step/conf1': _np(conf1), 'step/conf2': _np(conf2), 'step/cost1': _np(cost1), 'step/cost2': _np(cost2),
                'step/weights': np.array([1.0, 0.5, 1.0, 2.0], dtype=np.float32), 'step/min_conf_thr': np.array(10.0, dtype=np.float32),
                'step/loss': _np(loss), 'step/parts': np.array([me.batch_metrics[k][0] for k in
                                                                ('ap_loss', 'depth_loss', 'intra_depth_loss', 'kl_loss')]),
                'step/grad_layers': _np(torch.stack([torch.stack([t.grad[0, 1:] for t in tokens[v][0]]) for v in range(2)])),
                'step/grad_final': _np(torch.stack([tokens[v][1].grad[0, 1:] for v in range(2)])),
                'step/grad_head': _np(torch.cat([p.grad.reshape(-1) for p in (fl[0].weight, fl[0].bias, fl[1].weight,
                                                                              fl[1].bias, fl[3].weight, fl[3].bias)]))})
    print('training_step loss', float(loss), 'parts (ap, depth, intra, kl)', [round(me.batch_metrics[k][0], 5) for k in
                                                                              ('ap_loss', 'depth_loss', 'intra_depth_loss', 'kl_loss')])
    # ---- the same for FinetuneVGGTTIMM.training_step (src/finetune_timm_vggt.py:577-639): KL on block 7 only with
    #      co-visibility pixel masks, the vggt Smooth-AP variant, all loss weights 1.  Stand-ins: the teacher
    #      (extract_vggt_features, sample_keypoints -- its tracker picks the keypoints) and the ViT ----
    g = torch.Generator().manual_seed(950)
    vtokens = [([t.detach().clone().requires_grad_(True) for t in tokens[v][0]],
                tokens[v][1].detach().clone().requires_grad_(True)) for v in range(2)]
    vkp1 = synth.keypoints(951, K, mw, mh)[None]
    vkp2 = (vkp1 + torch.randint(-4, 5, vkp1.shape, generator=g).float())
    vkp2[..., 0].clamp_(3, mw - 4)
    vkp2[..., 1].clamp_(3, mh - 4)
    pm1 = (torch.rand(ph, pw, generator=g) < 0.6).repeat_interleave(14, 0).repeat_interleave(14, 1)
    pm2 = (torch.rand(ph, pw, generator=g) < 0.6).repeat_interleave(14, 0).repeat_interleave(14, 1)
    vcost1, vcost2 = synth.teacher_volume(952, N, 'vggt'), synth.teacher_volume(953, N, 'vggt')
    vfeat = dict(depth_pred_1=z1, depth_pred_2=z2, cost_1=vcost1[None], cost_2=vcost2[None], point_map_view_1=pts1,
                 point_map_view_2=pts2, image_shape=(mh, mw))
    vcls = ft_vggt.FinetuneVGGTTIMM
    vme = _Self({})
    vme.model = _PairViT(vtokens)
    vme.input_transform = lambda x: x
    vme.refine_conv = torch.nn.Identity()
    vme.target_res, vme.downsample_factor = max(mh, mw), 14
    vme.ap_loss_weight, vme.depth_loss_weight, vme.intra_depth_loss_weight, vme.kl_loss_weight = 1.0, 1.0, 1.0, 1.0
    torch.manual_seed(79)
    vme.depth_diff_head = ref_model.DepthAwareFeatureFusion(C)
    synth.load_head(vme.depth_diff_head, synth.head_params(4401, C))
    vme.log = lambda *a, **k: None
    vme.extract_vggt_features = lambda rgb, batch_idx=None: vfeat
    vme.sample_keypoints = lambda feats, num_keypoints=300, min_distance=5: (vkp1, vkp2, None, pm1, pm2)
    for name in ('calculate_depth_loss', 'calculate_cost_loss', 'calculate_matching_loss', 'get_intermediate_feature',
                 'get_feature', 'get_feature_cost'):
        setattr(vme, name, types.MethodType(getattr(vcls, name), vme))
    vbatch = dict(rgb_1=torch.zeros(1, 3, mh, mw), rgb_2=torch.ones(1, 3, mh, mw), rgb_vggt=None)
    vloss = vcls.training_step(vme, vbatch, 0)
    vloss.backward()
    fl = vme.depth_diff_head.fusion_layer
    out.update({'vstep/kp1': _np(vkp1), 'vstep/kp2': _np(vkp2), 'vstep/pixmask1': _np(pm1), 'vstep/pixmask2': _np(pm2),
                'vstep/cost1': _np(vcost1), 'vstep/cost2': _np(vcost2), 'vstep/loss': _np(vloss),
                'vstep/parts': np.array([vme.batch_metrics[k][0] for k in ('ap_loss', 'depth_loss', 'intra_depth_loss', 'kl_loss')]),
                'vstep/grad_layers': _np(torch.stack([torch.stack([t.grad[0, 1:] if t.grad is not None else torch.zeros_like(t[0, 1:])
                                                                   for t in vtokens[v][0]]) for v in range(2)])),
                'vstep/grad_final': _np(torch.stack([vtokens[v][1].grad[0, 1:] for v in range(2)])),
                'vstep/grad_head': _np(torch.cat([p.grad.reshape(-1) for p in (fl[0].weight, fl[0].bias, fl[1].weight,
                                                                               fl[1].bias, fl[3].weight, fl[3].bias)]))})
    print('vggt training_step loss', float(vloss), 'parts (ap, depth, intra, kl)',
          [round(vme.batch_metrics[k][0], 5) for k in ('ap_loss', 'depth_loss', 'intra_depth_loss', 'kl_loss')])
    out['meta'] = np.array([ph, pw, C, K])
    np.savez_compressed(os.path.join(OUT, 'live_bodies.npz'), **out)


def semantic_transfer_golden():
    """``tests/golden/eval_argmax.npz`` from the LIVE ``semantic_transfer`` (``src/evaluate_timm.py:461-588``).

    The function is run unmodified on CPU for one image pair.  Stand-ins: the dataset loader (two blank images and fixed
    keypoints), the ViT (fixed token tensors per image) and ``Tensor.cuda`` (identity: there is no GPU here);
    ``torch.argmax`` is wrapped for the duration of the call to record the similarity maxima and ``nn_idx``, which the
    function itself only folds into PCK numbers."""
    from PIL import Image
    import_live_modules()
    import evaluate_timm as ev
    img_size, ph, C, K = 640, 40, 24, 12
    g = torch.Generator().manual_seed(5)
    tok = [torch.randn(1, 1 + ph * ph, C, generator=g) for _ in range(2)]
    # image 2 = shifted, noisy copy of image 1, smoothed like real ViT features
    shifted = tok[0][:, 1:].reshape(1, ph, ph, C).roll(shifts=(2, -1), dims=(1, 2))
    tok[1][:, 1:] = (shifted + 0.3 * torch.randn(1, ph, ph, C, generator=g)).reshape(1, -1, C)
    kps = torch.zeros(2, K, 3)
    kps[:, :, :2] = torch.randint(20, 620, (2, K, 2), generator=g).float()
    kps[:, :, 2] = 1

    class _EvalViT:
        def forward_features(self, x):
            return tok[int(float(x.mean()) > 0)]          # image 1 is black, image 2 white (after imagenet_norm: < 0 / > 0)

    model = types.SimpleNamespace(model=_EvalViT())
    blank = {'img_a': Image.fromarray(np.zeros((480, 640, 3), np.uint8)),
             'img_b': Image.fromarray(np.full((480, 640, 3), 255, np.uint8))}
    seen = {}
    real = dict(argmax=torch.argmax, cuda=torch.Tensor.cuda, seed=torch.cuda.manual_seed, load=ev.load_pascal_data,
                open=ev.Image.open)

    def recording_argmax(x, *a, **k):
        res = real['argmax'](x, *a, **k)
        seen['best'], seen['nn_idx'] = x.max(dim=1).values.clone(), res.clone()
        return res
    try:
        torch.argmax = recording_argmax
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.manual_seed = lambda s: None
        ev.load_pascal_data = lambda *a, **k: (['img_a', 'img_b'], kps, None)
        ev.Image.open = lambda fn: blank[fn]
        ev.semantic_transfer(model, num_cats=1)
    finally:
        torch.argmax, torch.Tensor.cuda, torch.cuda.manual_seed = real['argmax'], real['cuda'], real['seed']
        ev.load_pascal_data, ev.Image.open = real['load'], real['open']
    out = {'tokens1': _np(tok[0][0, 1:]), 'tokens2': _np(tok[1][0, 1:]), 'kps1': _np(kps[0]),
           'nn_idx': _np(seen['nn_idx']), 'best': _np(seen['best']),
           'meta': np.array([img_size, 16, 16, ph, C, K])}
    np.savez_compressed(os.path.join(OUT, 'eval_argmax.npz'), **out)
    print('semantic_transfer nn_idx', seen['nn_idx'][:6].tolist())


if __name__ == '__main__':
    if '--eval' in sys.argv:
        semantic_transfer_golden()
    else:
        main()
