"""cfg3 of BASELINE.json: reciprocal NN on 8192 x 8192 x 24 descriptors (+ the real MASt3R shape).
Prints one JSON line with GPU kernel time, end-to-end call time and the CPU oracle time."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import numpy as np
import torch
from gd3 import _lib
from gd3.compat import fast_nn
from oracle import fast_nn as onn, synth


def main():
    out = {}
    A = synth.nn_exact_set(301, 8192); B = synth.nn_exact_set(302, 8192)
    Ad, Bd = A.cuda(), B.cuda()
    for _ in range(3): _lib.reciprocal_nn(Ad, Bd, dist='dot')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): nnA, nnB = _lib.reciprocal_nn(Ad, Bd, dist='dot')
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    for _ in range(5): a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist='dot', block_size=2 ** 13)
    call_ms = (time.perf_counter() - t0) / 5 * 1e3
    torch.set_num_threads(os.cpu_count())
    onn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot', block_size=2 ** 13)
    t0 = time.perf_counter()
    for _ in range(3): ra, rb = onn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot', block_size=2 ** 13)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    out['cfg3'] = dict(shape='8192x8192x24 dot', gpu_kernel_ms=round(ms, 4), gflops_fp32=round(2 * 8192 ** 2 * 24 / ms / 1e6, 1),
                       call_ms_host_in_numpy_out=round(call_ms, 3), cpu_oracle_ms=round(cpu_ms, 2), cpu_threads=torch.get_num_threads(),
                       bit_exact=bool((a == ra).all() and (b == rb).all()))
    d1, d2 = synth.nn_desc_maps(305, 384, 512)
    d1 = torch.round(d1 * 16) / 8; d2 = torch.round(d2 * 16) / 8
    d1c, d2c = d1.cuda(), d2.cuda()
    fast_nn.fast_reciprocal_NNs(d1c, d2c, subsample_or_initxy1=16, device='cuda', dist='dot', block_size=2 ** 13)
    t0 = time.perf_counter()
    for _ in range(5): xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1c, d2c, subsample_or_initxy1=16, device='cuda', dist='dot', block_size=2 ** 13)
    real_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    r1, r2 = onn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=16, device='cpu', dist='dot', block_size=2 ** 13)
    cpu_real_ms = (time.perf_counter() - t0) * 1e3
    out['mast3r_shape'] = dict(shape='384x512x24, S=16', matches=int(len(xy1)), call_ms=round(real_ms, 3), cpu_oracle_ms=round(cpu_real_ms, 1),
                               identical=bool(xy1.shape == r1.shape and (xy1 == r1).all() and (xy2 == r2).all()))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
