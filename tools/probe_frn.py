"""Per-kernel CUDA-event times of one MASt3R-shaped fast_reciprocal_NNs call (dev probe)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib
from gd3.compat import fast_nn
from oracle import synth

d1, d2 = synth.nn_desc_maps(305, 384, 512)
d1 = (torch.round(d1 * 16) / 8).cuda(); d2 = (torch.round(d2 * 16) / 8).cuda()
for S in (16, 8):
    for _ in range(2):
        fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=S, device='cuda', dist='dot')
    _lib.profile_enable(True); _lib.profile_read()
    t0 = time.perf_counter()
    xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=S, device='cuda', dist='dot')
    dt = (time.perf_counter() - t0) * 1e3
    prof = _lib.profile_read(); _lib.profile_enable(False)
    print(f'S={S}: {len(xy1)} matches, call {dt:.3f} ms (with per-kernel events)')
    for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print(f'   {k:20s} x{c:3d}  {ms * 1e3:8.1f} us total')
    t0 = time.perf_counter()
    for _ in range(5):
        fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=S, device='cuda', dist='dot')
    print(f'   call without events: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms')
