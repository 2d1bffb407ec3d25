"""One forward+backward of the cost-volume KL op at cfg2 size (target for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import ops

N, C, P = 1024, 768, 32
g = torch.Generator(device='cuda').manual_seed(0)
f1 = torch.randn(P, N, C, device='cuda', generator=g).to(torch.bfloat16).requires_grad_(True)
f2 = torch.randn(P, N, C, device='cuda', generator=g).to(torch.bfloat16).requires_grad_(True)
t12 = torch.softmax(4 * torch.randn(P, N, N, device='cuda', generator=g), -1)
t21 = torch.softmax(4 * torch.randn(P, N, N, device='cuda', generator=g), -1)
m1 = torch.rand(P, N, device='cuda', generator=g) < 0.6
m2 = torch.rand(P, N, device='cuda', generator=g) < 0.6
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    ops.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant='mast3r')
torch.cuda.synchronize()
print('ok')
