import sys; sys.path.insert(0,'/root/repo/3d-vlm-gd_b200')
import torch
from gd3 import _lib
def t(b, M, N, K, tile_n, mn, iters=20):
    A = torch.randn(b, M, K, device='cuda').to(torch.bfloat16)
    B = torch.randn(b, K, N, device='cuda').to(torch.bfloat16) if mn else torch.randn(b, N, K, device='cuda').to(torch.bfloat16)
    f = (lambda: _lib.debug_gemm_bf16_mn(A, B, a_mn=False, b_mn=True, tile_n=tile_n)) if mn else (lambda: _lib.debug_gemm_bf16(A, B, tile_n=tile_n))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for mn in (False, True):
    for tile_n in (128, 192, 256):
        us = t(32, 1024, 768, 1024, tile_n, mn)
        print(f'mn={mn} tile_n={tile_n}: {us:.1f} us ({2.0*32*1024*768*1024/us/1e6:.0f} TF)', flush=True)
