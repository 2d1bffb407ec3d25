"""Fused teacher cost-volume post-processing at the MASt3R size (12 layers x 12 heads, N = 768 / 1024)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib
from oracle import teacher as ot

out = {}
for N in (768, 1024):
    L = H = 12
    g = torch.Generator(device='cuda').manual_seed(N)
    tgt = [3 * torch.randn(1, H, N, N, device='cuda', generator=g) for _ in range(L)]
    src = [3 * torch.randn(1, H, N, N, device='cuda', generator=g) for _ in range(L)]
    for _ in range(3): o = _lib.teacher_volume(tgt, src, 3.0, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): o = _lib.teacher_volume(tgt, src, 3.0, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    in_bytes = 2.0 * L * H * N * N * 4
    rec = dict(N=N, layers=L, heads=H, gpu_ms=round(ms, 4), input_GB=round(in_bytes / 1e9, 3),
               GBps=round((in_bytes + 2 * L * N * N * 4 + N * N * 4) / ms / 1e6, 1))
    # the same torch ops on the GPU (what the reference executes), for scale
    for _ in range(2): r = ot.teacher_volume(tgt, src, 3.0, True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): r = ot.teacher_volume(tgt, src, 3.0, True)
    e1.record(); torch.cuda.synchronize()
    rec['torch_gpu_ms'] = round(e0.elapsed_time(e1) / 5, 3)
    rec['max_abs_diff_vs_torch_gpu'] = float((o - r).abs().max())
    if N == 768:
        torch.set_num_threads(os.cpu_count())
        tc, sc = [t.cpu() for t in tgt], [s.cpu() for s in src]
        t0 = time.perf_counter(); ot.teacher_volume(tc, sc, 3.0, True)
        rec['cpu_oracle_ms'] = round((time.perf_counter() - t0) * 1e3, 1); rec['cpu_threads'] = torch.get_num_threads()
    out[f'N{N}'] = rec
print(json.dumps(out))
