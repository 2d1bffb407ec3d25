"""Time the cost-volume KL pipeline alone (default cfg2: 32 pairs x 1024 tokens x 768), fp32 and packed teachers.
    python tools/time_kl.py [--N 1024] [--C 768] [--P 32] [--iters 10]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib, ops
ap = argparse.ArgumentParser()
ap.add_argument('--N', type=int, default=1024); ap.add_argument('--C', type=int, default=768)
ap.add_argument('--P', type=int, default=32); ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--variant', default='mast3r'); ap.add_argument('--tag', default='')
a = ap.parse_args()
g = torch.Generator(device='cuda'); g.manual_seed(1)
P, N, C = a.P, a.N, a.C
f1 = torch.randn(P, N, C, generator=g, device='cuda').to(torch.bfloat16)
f2 = torch.randn(P, N, C, generator=g, device='cuda').to(torch.bfloat16)
t = [torch.softmax(4 * torch.randn(P, N, N, generator=g, device='cuda'), -1) for _ in range(2)]
m1 = torch.rand(P, N, generator=g, device='cuda') < 0.6
m2 = torch.rand(P, N, generator=g, device='cuda') < 0.6
packed = [ops.pack_teacher(x) for x in t]
for name, tt in (('fp32 teacher', t), ('packed teacher', packed)):
    for _ in range(3):
        out = ops.cost_kl_raw(f1, f2, tt[0], tt[1], m1, m2, a.variant)
    torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_read()
    for _ in range(a.iters):
        out = ops.cost_kl_raw(f1, f2, tt[0], tt[1], m1, m2, a.variant)
    torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    tot = sum(ms for _, ms in prof.values())
    print(f'[{a.tag}] {name}: N={N} C={C} P={P}: pipeline {tot / a.iters * 1e3:.1f} us; ' +
          ', '.join(f'{k} {ms / a.iters * 1e3:.1f}' for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]),
          f'loss {float(out[0].double().mean()):.6f}')
