"""Random-shape sweep of the fused step against the CPU oracle (dev probe: hunts shape-specific bugs).

    python tools/probe_shapes.py [n_cases] [seed]
"""
import os
import random
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-vlm-gd_b200'), os.path.join(ROOT, 'tests')]

import bench_common                        # noqa: E402
from gd3 import pipeline                   # noqa: E402


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300))


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for case in range(n_cases):
        if os.environ.get('GD3_PROBE_LARGE'):
            # larger keypoint counts (Smooth-AP rows beyond one warp's registers, several ranking tiles) and token grids
            ph, pw = rnd.randint(20, 40), rnd.randint(20, 40)
            cfg = dict(N=ph * pw, C=rnd.choice([64, 104, 200]), K=rnd.choice([513, 700, 1025, 1100]),
                       grid=(ph, pw), P=rnd.randint(1, 2), variant=rnd.choice(['mast3r', 'vggt']))
        else:
            ph, pw = rnd.randint(4, 24), rnd.randint(4, 24)
            cfg = dict(N=ph * pw, C=rnd.choice([64, 72, 96, 100, 128, 200, 384, 388, 520]),
                       K=rnd.choice([1, 2, 7, 33, 64, 100, 129, 257]), grid=(ph, pw), P=rnd.randint(1, 4),
                       variant=rnd.choice(['mast3r', 'vggt']))
        dtype = rnd.choice([torch.float32, torch.bfloat16])
        tag = f"case {case}: {cfg} {dtype}"
        try:
            batch = bench_common.make_batch(cfg, cfg_id=1, pair0=case)
            want = bench_common.oracle_step(batch, cfg)
            out = pipeline.distillation_step(bench_common.to_device(batch, 'cuda', feature_dtype=dtype),
                                             variant=cfg['variant'], grid=cfg['grid'], pairs_per_group=rnd.choice([0, 1, 2]))
            torch.cuda.synchronize()
            msgs = []
            for k in ('kl', 'ap', 'rank', 'l1'):
                got, ref = out[k].float().cpu(), want[k]
                err = ((got - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
                if not err <= 1e-3:
                    msgs.append(f'{k} rel err {err:.2e} got {got.tolist()} ref {ref.tolist()}')
            for k in ('f1', 'f2', 'g1', 'g2', 'head'):
                g, r = out['grads'][k].float().cpu(), want['grads'][k]
                if float(r.norm()) < 1e-12 and float(g.norm()) < 1e-9:
                    continue
                c = cos(g, r)
                if not c >= 0.999:
                    msgs.append(f'grad {k} cos {c:.5f} norms {float(g.norm()):.3e} {float(r.norm()):.3e}')
            print(('FAIL ' if msgs else 'ok   ') + tag)
            for m in msgs:
                print('     ', m)
            bad += bool(msgs)
        except Exception:
            bad += 1
            print('EXC  ' + tag)
            traceback.print_exc()
    print('bad cases:', bad)


if __name__ == '__main__':
    main()
