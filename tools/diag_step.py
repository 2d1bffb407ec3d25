import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench_common
from gd3 import pipeline, ops
from helpers import cosine

for variant, cfg in [('vggt', dict(N=15 * 17, C=200, K=77, grid=(15, 17), P=2)), ('vggt', dict(N=256, C=384, K=128, grid=(16, 16), P=2)),
                     ('mast3r', dict(N=15 * 17, C=200, K=77, grid=(15, 17), P=2))]:
    cfg = dict(cfg, variant=variant)
    batch = bench_common.make_batch(cfg, cfg_id=1)
    for wts in (None, dict(ap=0, kl=0, intra=1, depth=0), dict(ap=0, kl=0, intra=0, depth=1)):
        if wts is not None:
            bench_common.WEIGHTS[variant] = dict(wts)
        want = bench_common.oracle_step(batch, cfg)
        dev = bench_common.to_device(batch, 'cuda')
        out = pipeline.distillation_step(dev, variant=variant, grid=cfg['grid'], backward=True, weights=wts)
        D = cfg['C']
        gp = ops.split_param_grads(out['grads']['head'].cpu(), D)
        wp = ops.split_param_grads(want['grads']['head'], D)
        print(variant, cfg['K'], wts, 'losses', {k: (out[k].cpu().tolist(), want[k].tolist()) for k in ('rank', 'l1')})
        print('   head cos', [round(cosine(a, b), 5) for a, b in zip(gp, wp)], 'norms', [(round(float(a.norm()), 4), round(float(b.norm()), 4)) for a, b in zip(gp, wp)])
        print('   g1 cos', cosine(out['grads']['g1'], want['grads']['g1']), 'g2', cosine(out['grads']['g2'], want['grads']['g2']))
