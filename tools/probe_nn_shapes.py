"""Random-size sweep of gd3_reciprocal_nn against the CPU oracle on exactly-representable descriptors (bit-exact
indices incl. ties).  Dev probe.    python tools/probe_nn_shapes.py [n_cases] [seed]"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-vlm-gd_b200')]

from gd3.compat import fast_nn             # noqa: E402
from oracle import fast_nn as oracle_nn    # noqa: E402
from oracle import synth                   # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for case in range(n_cases):
        na = rnd.choice([1, 2, 31, 64, 65, 127, 129, 500, 1000, 2049, 3000])
        nb = rnd.choice([1, 3, 33, 128, 130, 777, 1024, 2500, 4097])
        dim = rnd.choice([1, 2, 3, 5, 8, 24, 25, 64, 100, 128, 130, 256])
        dist = rnd.choice(['dot', 'l2'])
        A = synth.nn_exact_set(10 * case + 1, na, dim=dim, dup=min(8, na // 2))
        B = synth.nn_exact_set(10 * case + 2, nb, dim=dim, dup=min(8, nb // 2))
        a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist=dist)
        ra, rb = oracle_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist=dist)
        ok = (a == ra).all() and (b == rb).all()
        print(('ok   ' if ok else 'FAIL ') + f'case {case}: nA={na} nB={nb} dim={dim} {dist}'
              + ('' if ok else f' mism A {int((a != ra).sum())} B {int((b != rb).sum())}'))
        bad += not ok
    print('bad cases:', bad)


if __name__ == '__main__':
    main()
