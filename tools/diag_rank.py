import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, copy
from oracle import bodies, losses as olosses, synth
from gd3 import ops
from helpers import cosine

def run(K, D, P, l1w, seed=0):
    head = olosses.DepthHead(D); synth.load_head(head, synth.head_params(81 + seed, D))
    gen = synth._gen(82 + seed)
    feats = 0.5 * torch.randn(2 * P, K, D, generator=gen)
    depths = torch.stack([synth.depths(90 + s, K) for s in range(2 * P)])
    ref_f = feats.clone().requires_grad_(True)
    tot = 0.0
    for p in range(P):
        l1, rk = bodies.depth_losses(head, ref_f[2*p:2*p+1], ref_f[2*p+1:2*p+2], depths[2*p:2*p+1], depths[2*p+1:2*p+2])
        tot = tot + rk + l1w * l1
    tot.backward()
    fl = head.fusion_layer
    want = [q.grad.clone() for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight, fl[3].bias)]
    h = copy.deepcopy(head).cuda()
    for prm in h.parameters(): prm.grad = None
    x = feats.cuda().requires_grad_(True)
    total, lr, l1 = ops.depth_head_loss(h, x, depths.cuda(), w_rank=torch.full((2*P,), 0.5, device='cuda'),
                                        w_l1=torch.full((P,), l1w, device='cuda') if l1w is not None else None)
    total.backward()
    fl = h.fusion_layer
    got = [q.grad for q in (fl[0].weight, fl[0].bias, fl[1].weight, fl[1].bias, fl[3].weight, fl[3].bias)]
    names = ['W1','b1','gamma','beta','w2','b2']
    print(f'K={K} D={D} P={P} l1w={l1w}: total {total.item():.6f} vs {float(tot):.6f}; feats cos {cosine(x.grad, ref_f.grad):.6f}; ' +
          ' '.join(f'{n}:{cosine(g, w):.5f}({float(g.norm()):.3g}/{float(w.norm()):.3g})' for n, g, w in zip(names, got, want)))

for K, D, P, l1w in [(77, 200, 2, 1.0), (77, 200, 2, 0.0), (80, 200, 2, 1.0), (77, 256, 2, 1.0), (128, 384, 2, 1.0), (128, 384, 3, 0.33), (48, 64, 1, 1.0), (77, 200, 1, 1.0)]:
    run(K, D, P, l1w)
