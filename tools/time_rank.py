"""Time the depth-ranking pipeline alone at a BASELINE shape (default cfg2: 64 sets x 512 keypoints x 768) and print a
signature of the losses / gradients (correctness is the job of tests/test_gpu_depth_rank.py).

    python tools/time_rank.py [--K 512] [--D 768] [--S 64] [--iters 10]
Prints per-kernel CUDA-event times (gd3_profile_*), rank_pairs first."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument('--K', type=int, default=512)
ap.add_argument('--D', type=int, default=768)
ap.add_argument('--S', type=int, default=64)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--tag', default='')
a = ap.parse_args()
g = torch.Generator().manual_seed(5)
S, K, D = a.S, a.K, a.D
feats = (0.5 * torch.randn(S, K, D, generator=g)).cuda()
depths = (torch.rand(S, K, generator=g) * 4.5 + 0.5).cuda()
k1, k2 = D ** -0.5, 128 ** -0.5
params = [((torch.rand(128, D, generator=g) * 2 - 1) * k1), ((torch.rand(128, generator=g) * 2 - 1) * k1),
          1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g),
          ((torch.rand(1, 128, generator=g) * 2 - 1) * k2), ((torch.rand(1, generator=g) * 2 - 1) * k2)]
params = [p.cuda() for p in params]
w_rank = torch.full((S,), 0.5 / (S // 2), device='cuda')
w_l1 = torch.full((S // 2,), 1.0 / (S // 2), device='cuda')


def run():
    return ops.depth_head_raw(feats, depths, params, True, 1e-5, 0, 0.05, 0.05, False, w_rank, w_l1, True)


for _ in range(3):
    out = run()
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_read()
for _ in range(a.iters):
    out = run()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
tot = sum(ms for _, ms in prof.values())
print(f'[{a.tag}] S={S} K={K} D={D}: pipeline {tot / a.iters * 1e3:.1f} us; ' +
      ', '.join(f'{k} {ms / a.iters * 1e3:.1f}' for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:16]))
lr, l1, gf, gp = out
sig = dict(loss=float(lr.double().mean()), l1=float(l1.double().mean()), gf=float(gf.double().norm()), gp=float(gp.double().norm()))
print("  ", sig, flush=True)
