"""Throughput probe of the tcgen05 GEMM variants (dev tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib

def bench(b, M, N, K, tile_n, iters=20):
    A = torch.randn(b, M, K, device='cuda').to(torch.bfloat16)
    B = torch.randn(b, N, K, device='cuda').to(torch.bfloat16)
    for _ in range(3): C = _lib.debug_gemm_bf16(A, B, tile_n=tile_n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): C = _lib.debug_gemm_bf16(A, B, tile_n=tile_n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ref = torch.matmul(A[0].float(), B[0].float().t())
    err = (C[0] - ref).abs().max().item()
    print(f'b={b} M={M} N={N} K={K} tile_n={tile_n}: {ms*1e3:.1f} us  {2.0*b*M*N*K/ms/1e9:.0f} TFLOP/s  maxerr {err:.3g}', flush=True)

if __name__ == '__main__':
    for tile_n in (256, -256, 128, -128):
        bench(32, 1024, 768, 1024, tile_n)      # KL gradient GEMM, cfg2
        bench(32, 1024, 1024, 768, tile_n)      # KL z GEMM, cfg2
        bench(1, 8192, 8192, 8192, tile_n, iters=5)
    A = torch.randn(8192, 8192, device='cuda').to(torch.bfloat16); B = torch.randn(8192, 8192, device='cuda').to(torch.bfloat16)
    for _ in range(3): A @ B.t()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): A @ B.t()
    e1.record(); torch.cuda.synchronize()
    print(f'cuBLAS 8192^3: {2*8192**3/(e0.elapsed_time(e1)/5)/1e9:.0f} TFLOP/s')


def cublas_bmm(b, M, N, K, iters=20):
    A = torch.randn(b, M, K, device='cuda').to(torch.bfloat16)
    B = torch.randn(b, N, K, device='cuda').to(torch.bfloat16)
    for _ in range(3): torch.bmm(A, B.transpose(1, 2))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): torch.bmm(A, B.transpose(1, 2))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'cuBLAS bmm b={b} M={M} N={N} K={K}: {ms*1e3:.1f} us  {2.0*b*M*N*K/ms/1e9:.0f} TFLOP/s (bf16 out)', flush=True)


if __name__ == '__main__':
    cublas_bmm(32, 1024, 768, 1024)
    cublas_bmm(32, 1024, 1024, 768)
    cublas_bmm(64, 1369, 1024, 1369)
