"""Quick timing probe for the cost-volume KL op (dev tool, not part of the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import ops, _lib

def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    N, C, P = (1024, 768, 32) if cfg == 'cfg2' else (1369, 1024, 64)
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    f1 = torch.randn(P, N, C, device=dev, generator=g).to(torch.bfloat16)
    f2 = torch.randn(P, N, C, device=dev, generator=g).to(torch.bfloat16)
    t12 = torch.softmax(4 * torch.randn(P, N, N, device=dev, generator=g), -1)
    t21 = torch.softmax(4 * torch.randn(P, N, N, device=dev, generator=g), -1)
    m1 = torch.rand(P, N, device=dev, generator=g) < 0.6
    m2 = torch.rand(P, N, device=dev, generator=g) < 0.6
    lib = _lib.load()
    for ppg in [0, 2, 4, 6, 8, 16, 32, 64]:
        if ppg > P: continue
        G = lib.gd3_cost_kl_group_size(P, N, C, ppg)
        f1.requires_grad_(True); f2.requires_grad_(True)
        for _ in range(3):
            loss = ops.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant='mast3r', pairs_per_group=ppg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            loss = ops.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant='mast3r', pairs_per_group=ppg)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 6.0 * N * N * C * P
        print(f'{cfg} ppg={ppg} G={G} fwd+bwd {ms:.3f} ms  {P / ms * 1e3:.0f} pairs/s  {flops / ms / 1e9:.1f} TFLOP/s algorithmic', flush=True)
        with torch.no_grad():
            for _ in range(2):
                ops.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant='mast3r', pairs_per_group=ppg)
            torch.cuda.synchronize(); e0.record()
            for _ in range(iters):
                ops.cost_volume_kl(f1, f2, t12, t21, m1, m2, variant='mast3r', pairs_per_group=ppg)
            e1.record(); torch.cuda.synchronize()
        print(f'   fwd only {e0.elapsed_time(e1) / iters:.3f} ms', flush=True)

if __name__ == '__main__':
    main()
