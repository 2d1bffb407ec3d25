"""Reciprocal-NN calls at cfg3 (8192 x 8192 x 24, dot) on device-resident descriptors: the ncu target for nn_tile_kernel
(`one_nn.py 3`) and a timing loop (`one_nn.py 20`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib
from oracle import synth

A = synth.nn_exact_set(301, 8192).cuda(); B = synth.nn_exact_set(302, 8192).cuda()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ref = None
for shift in (sys.argv[2:] or [None]):
    if shift is not None:
        os.environ['GD3_NN_SHIFT'] = shift
    for _ in range(iters):
        out = _lib.reciprocal_nn(A, B, dist='dot')
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    assert all(bool((a == b).all()) for a, b in zip(out, ref))
    if iters > 3:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): _lib.reciprocal_nn(A, B, dist='dot')
        e1.record(); torch.cuda.synchronize()
        print('shift', shift, 'ms per call', round(e0.elapsed_time(e1) / 100, 4))
