"""Where the step time goes at the strong-scaling shard sizes of cfg4 (64 pairs over 8 / 4 GPUs = 8 / 16 pairs per GPU):
graph replay time with and without parallel branches and the eager per-kernel times, on ONE GPU.
    python tools/diag_strong.py [P ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
import bench_extras
from bench import WORKLOADS
from gd3 import _lib, pipeline

dev = torch.device('cuda', 0)
c = WORKLOADS['cfg4']
N, C, K, grid, variant = c['N'], c['C'], c['K'], c['grid'], c['variant']
for P in [int(a) for a in sys.argv[1:]] or [8, 16, 64]:
    batch = bench_extras.make_device_batch(P, N, C, K, grid, variant, dev)
    res = {}
    for par in (True, False):
        gs = pipeline.GraphedStep(batch, variant=variant, grid=grid, backward=True, parallel_branches=par)
        for _ in range(5): gs()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): gs()
        e1.record(); torch.cuda.synchronize()
        res[par] = e0.elapsed_time(e1) / 30
        del gs
    _lib.profile_enable(True); _lib.profile_read()
    for _ in range(3):
        pipeline.distillation_step(batch, variant=variant, grid=grid, backward=True)
    prof = _lib.profile_read(); _lib.profile_enable(False)
    tot = sum(ms for _, ms in prof.values()) / 3
    print(f'P={P}: graph parallel {res[True]:.4f} ms, single stream {res[False]:.4f} ms, sum of kernels {tot:.4f} ms')
    for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print(f'    {k:24s} {cnt // 3:3d} x {ms / cnt * 1e3:8.1f} us = {ms / 3 * 1e3:8.1f} us')
    del batch
