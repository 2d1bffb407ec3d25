import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/3d-vlm-gd_b200')
import torch
from gd3 import ops
from oracle import bodies, synth
torch.manual_seed(0)
N, C = 256, 384
for spread in (1.0, 0.3, 0.1, 0.03):
    base = torch.randn(1, 1, C)
    f1 = (base + spread * torch.randn(1, N, C)).to(torch.bfloat16)
    f2 = (base + spread * torch.randn(1, N, C)).to(torch.bfloat16)
    t12 = synth.teacher_volume(1, N, 'mast3r')[None]; t21 = synth.teacher_volume(2, N, 'mast3r')[None]
    m1 = torch.ones(1, N, dtype=torch.bool); m2 = torch.ones(1, N, dtype=torch.bool)
    a = f1.float().requires_grad_(True); b = f2.float().requires_grad_(True)
    want = bodies.cost_volume_kl(a[0], b[0], t12[0], t21[0], m1[0], m2[0], variant='mast3r')
    want.backward()
    x = f1.cuda().requires_grad_(True); y = f2.cuda().requires_grad_(True)
    got = ops.cost_volume_kl(x, y, t12.cuda(), t21.cuda(), m1.cuda(), m2.cuda(), variant='mast3r')
    got.sum().backward()
    cos = torch.nn.functional.cosine_similarity(x.grad.float().cpu().flatten(), a.grad.flatten(), dim=0).item()
    print(f'spread {spread}: loss got {got.item():.6f} want {want.item():.6f} rel {abs(got.item()-want.item())/abs(want.item()):.2e} grad cos {cos:.5f}')
