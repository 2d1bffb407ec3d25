"""Time the Smooth-AP pipeline alone (default cfg2: 32 pairs x 512 keypoints x 768)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib, ops
ap = argparse.ArgumentParser()
ap.add_argument('--K', type=int, default=512); ap.add_argument('--C', type=int, default=768)
ap.add_argument('--P', type=int, default=32); ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--variant', default='mast3r'); ap.add_argument('--tag', default='')
a = ap.parse_args()
g = torch.Generator(device='cuda'); g.manual_seed(2)
P, K, C = a.P, a.K, a.C
base = torch.randn(1, 1, C, generator=g, device='cuda')
d1 = torch.nn.functional.normalize(base + 0.12 * torch.randn(P, K, C, generator=g, device='cuda'), dim=-1)
d2 = torch.nn.functional.normalize(d1 + 0.03 * torch.randn(P, K, C, generator=g, device='cuda'), dim=-1)
p1 = torch.rand(P, K, 3, generator=g, device='cuda'); p2 = p1 + 0.02 * torch.randn(P, K, 3, generator=g, device='cuda')
for _ in range(3):
    out = ops.smooth_ap_raw(d1, d2, p1, p2, a.variant)
torch.cuda.synchronize()
_lib.profile_enable(True); _lib.profile_read()
for _ in range(a.iters):
    out = ops.smooth_ap_raw(d1, d2, p1, p2, a.variant)
torch.cuda.synchronize()
prof = _lib.profile_read(); _lib.profile_enable(False)
tot = sum(ms for _, ms in prof.values())
print(f'[{a.tag}] K={K} C={C} P={P}: pipeline {tot / a.iters * 1e3:.1f} us; ' +
      ', '.join(f'{k} {ms / a.iters * 1e3:.1f}' for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]),
      f'loss {float(out[0].double().mean()):.6f}')
