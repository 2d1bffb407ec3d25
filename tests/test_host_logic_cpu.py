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
# uneven shards: rank gradients are means over P_local pairs; the weighted form is the mean over all P pairs
gw = torch.full((3,), float(sum(range(b, e))) / max(e - b, 1))
gdist.allreduce_mean_(gw, local_pairs=e - b, num_pairs=P)
assert torch.allclose(gw, torch.full((3,), sum(range(P)) / P)), gw
# sharding goes by key: per-pair tensors are sliced, a shared (H, W) depth map with H == P is not
batch = dict(f1=torch.arange(P * 2.).reshape(P, 2), kp1=torch.zeros(P, 3, 2), depth_map1=torch.ones(P, 5),
             h1=torch.zeros(4, P, 6, 2), head=dict(W1=torch.zeros(P, 2)), scale=torch.ones(P))
sh = gdist.shard_batch(batch, rank, world)
assert sh['f1'].shape[0] == e - b and sh['kp1'].shape[0] == e - b and sh['h1'].shape[:2] == (4, e - b)
assert sh['depth_map1'].shape == (P, 5) and sh['scale'].shape == (P,) and sh['head']['W1'].shape == (P, 2)
# bucketed, overlapped gradient all-reduce (the DDP protocol of src/main.py:147-159) against plain averaging
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 16), torch.nn.Linear(16, 1))
net[2].weight.requires_grad_(False)                         # a frozen parameter is never bucketed
unused = torch.nn.Parameter(torch.ones(7))                   # a trainable parameter that gets no gradient
params = list(net.parameters()) + [unused]
red = gdist.BucketedGradAllReduce(params, bucket_bytes=100)  # tiny buckets -> several collectives
assert len(red.buckets) > 2
for step in range(2):
    x = torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * step + rank))
    for q in params:
        q.grad = None
    red.reset()
    (net(x).pow(2).mean() * (rank + 1)).backward()
    local = [None if q.grad is None else q.grad.clone() for q in params]
    norm = red.finish(clip_norm=0.05)
    want = []
    for q, gl in zip(params, local):
        if not q.requires_grad:
            continue
        t = torch.zeros_like(q) if gl is None else gl.clone()
        dist.all_reduce(t)
        want.append((q, t / world))
    tot = torch.sqrt(sum((t ** 2).sum() for _, t in want))
    sc = min(1.0, 0.05 / (float(tot) + 1e-6))
    assert abs(float(norm) - float(tot)) < 1e-5 * float(tot)
    for q, t in want:
        assert torch.allclose(q.grad, t * sc, rtol=1e-5, atol=1e-7)
red.remove()
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
