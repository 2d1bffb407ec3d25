"""Separate per-tile fixed cost from per-k-block cost of the tcgen05 GEMM variants: time(K) = a + b*K."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib

def t(b, M, N, K, tile_n, iters=20):
    A = torch.randn(b, M, K, device='cuda').to(torch.bfloat16)
    B = torch.randn(b, N, K, device='cuda').to(torch.bfloat16)
    for _ in range(3): _lib.debug_gemm_bf16(A, B, tile_n=tile_n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): _lib.debug_gemm_bf16(A, B, tile_n=tile_n)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for tile_n in (256, 192, -256, -192):
    row = []
    for K in (256, 1024, 2048, 4096):
        us = t(32, 1024, 768, K, tile_n)
        row.append(f'K={K}: {us:.1f} us ({2.0*32*1024*768*K/us/1e6:.0f} TF)')
    print(f'tile_n={tile_n}:', '  '.join(row), flush=True)
