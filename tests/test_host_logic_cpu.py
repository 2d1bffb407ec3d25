"""CPU-only tests of the host-side logic: numpy bookkeeping of the fast_nn drop-in, pair sharding and the
world_size-2 gloo path of gd3.dist."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_merge_corres_matches_reference_golden(golden):
    from gd3.compat import fast_nn
    g = golden('fast_nn.npz')
    x1, x2 = fast_nn.merge_corres(g['merge/i1'], g['merge/i2'], (48, 64), (48, 64))
    assert (x1 == g['merge/xy1']).all() and (x2 == g['merge/xy2']).all()
    assert x1.dtype == g['merge/xy1'].dtype
    j1, j2, jdx = fast_nn.merge_corres(g['merge/i1'], g['merge/i2'], ret_xy=False, ret_index=True)
    assert (j1 == g['merge/j1']).all() and (j2 == g['merge/j2']).all() and (jdx == g['merge/jdx']).all()
    (y1, xx1), _ = fast_nn.merge_corres(g['merge/i1'], g['merge/i2'], (48, 64), (48, 64), ret_xy='y_x')
    assert (np.stack([xx1, y1], -1) == g['merge/xy1']).all()
    with pytest.raises(AssertionError):
        fast_nn.merge_corres(g['merge/i1'].astype(np.int64), g['merge/i2'].astype(np.int64))


def test_fast_nn_reference_error_behaviour():
    from gd3 import _lib
    from gd3.compat import fast_nn
    a = torch.randn(4, 5, 3)
    with pytest.raises(AssertionError):                       # DIM1 == DIM2 (mast3r/fast_nn.py:113)
        fast_nn.fast_reciprocal_NNs(a, torch.randn(4, 5, 2), device='cuda')
    with pytest.raises(_lib.Gd3Error):                        # KDTree CPU branch is not provided
        fast_nn.fast_reciprocal_NNs(a, a, device='cpu')
    with pytest.raises(ValueError):                           # Unknown dist (mast3r/fast_nn.py:37)
        fast_nn.bruteforce_reciprocal_nns(a[0], a[0], device='cuda', dist='cosine')
    m = fast_nn.cdistMatcher.__new__(fast_nn.cdistMatcher)
    with pytest.raises(AssertionError):                       # assert k == 1 (mast3r/fast_nn.py:79)
        m.query(torch.zeros(1, 3), k=2)
    assert m.query(torch.zeros(0, 3)) == (None, [])           # empty query (:80-81)


def test_shard_range_partitions_exactly():
    from gd3 import dist as gdist
    for P in (0, 1, 7, 32, 64, 65):
        for W in (1, 2, 3, 4, 8):
            spans = [gdist.shard_range(P, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        gdist.shard_range(8, 2, 2)
    batch = dict(f1=torch.arange(10)[:, None, None].float(), kp1=torch.zeros(10, 3, 2), head=dict(W1=torch.zeros(2, 2)))
    part = gdist.shard_batch(batch, 1, 4)
    assert part['f1'].flatten().tolist() == [3.0, 4.0, 5.0] and part['head'] is batch['head']


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], '3d-vlm-gd_b200'))
from gd3 import dist as gdist
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
P = 7
b, e = gdist.shard_range(P, rank, world)
# stand-in for the per-pair loss op: a deterministic function of the global pair index
local = torch.arange(b, e, dtype=torch.float32) * 1.5 + 0.25
full = gdist.gather_pair_losses(local, P)
assert torch.equal(full, torch.arange(P, dtype=torch.float32) * 1.5 + 0.25), full
g = torch.full((5,), float(rank + 1))
gdist.allreduce_mean_(g)
assert torch.allclose(g, torch.full((5,), sum(range(1, world + 1)) / world))
dist.barrier()
if rank == 0:
    print('GLOO_OK')
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29613', str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'GLOO_OK' in r.stdout
