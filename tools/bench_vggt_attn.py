"""Time gd3_vggt_attn_accumulate at VGGT's shape against the same torch ops on the same GPU (dev probe).

    python tools/bench_vggt_attn.py            # n = 925 (25 x 37 patches), 16 heads, head_dim 64, 24 blocks
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-vlm-gd_b200')]

from gd3 import _lib                      # noqa: E402
from gd3.compat import teacher            # noqa: E402
from oracle import synth                  # noqa: E402


def torch_block(q, k, scale, temp, skip=5):
    """The reference's ops (vggt/layers/attention.py:73-84) under CUDA bf16 autocast."""
    N = q.shape[-2]
    with torch.autocast('cuda', dtype=torch.bfloat16):
        qs = q * scale
        s1 = torch.matmul(qs[..., skip:N // 2, :], k[..., N // 2 + skip:, :].transpose(-2, -1))
        a1 = torch.softmax(s1 / temp, dim=-1)
        s2 = torch.matmul(qs[..., N // 2 + skip:, :], k[..., skip:N // 2, :].transpose(-2, -1))
        a2 = torch.softmax(s2 / temp, dim=-1)
    return torch.cat([a1, a2], dim=0)


def main():
    n, heads, blocks = 925, 16, 24
    qk = [tuple(t.cuda() for t in synth.vggt_qk(100 + b, 1, heads, n)) for b in range(blocks)]

    def ours():
        vols = teacher.VggtCostVolumes(blocks, 1.0)
        for q, k in qk:
            vols.add_block(q, k, 0.125)
        return vols.result()

    def ref():
        maps = [torch_block(q, k, 0.125, 1.0) for q, k in qk]
        attn = torch.mean(torch.stack(maps), dim=0)
        c1, c2 = attn.chunk(2, dim=0)
        return c1.mean(dim=1), c2.mean(dim=1)

    out = {}
    for name, fn in (('gd3', ours), ('torch_same_gpu', ref)):
        for _ in range(3):
            r = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        out[name + '_ms'] = round(e0.elapsed_time(e1) / 10, 3)
        out[name] = r
    d1 = float((out['gd3'][0] - out['torch_same_gpu'][0]).abs().sum() / out['torch_same_gpu'][0].abs().sum())
    print(json.dumps(dict(workload=f'VGGT cost volumes: n={n}, heads={heads}, head_dim=64, {blocks} blocks, B=1',
                          gd3_ms=out['gd3_ms'], torch_same_gpu_ms=out['torch_same_gpu_ms'],
                          rel_l1_diff_vs_torch=d1, peak_mem_note='torch path stacks 24 x (2, 16, 925, 925) fp32 maps = 2.6 GB')))


if __name__ == '__main__':
    main()
