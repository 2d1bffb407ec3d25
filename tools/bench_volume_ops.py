"""Timing of the volume-level helpers (gd3.compat get_masked_patch_cost / kl_divergence_map) against the same
expressions in eager PyTorch, on cfg2-sized volumes (32 x 1024 x 1024 fp32).  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib
from gd3.compat import functions as fn, losses


def timed(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    P, N = 32, 1024
    g = torch.Generator(device='cuda').manual_seed(1)
    t = torch.softmax(4 * torch.randn(P, N, N, device='cuda', generator=g), -1)
    z = torch.randn(P, N, N, device='cuda', generator=g)
    m1 = torch.rand(N, device='cuda', generator=g) < 0.6
    mb = P * N * N * 4 / 1e6

    def ours():
        zz = z.detach().requires_grad_(True)
        kl = losses.kl_divergence_map(fn.get_masked_patch_cost(t, m1), fn.get_masked_patch_cost(zz, m1, use_softmax=True))
        kl.backward()
        return zz.grad

    def eager():
        zz = z.detach().requires_grad_(True)
        def mpc(c, sm):
            o = torch.where(m1[None, :, None], c, c.new_zeros(()))
            return torch.softmax(o, -1, dtype=torch.float32) if sm else o / o.sum(-1, keepdim=True).clamp_min(1e-8)
        a, b = mpc(t, False).clamp_min(1e-8), mpc(zz, True).clamp_min(1e-8)
        kl = (a * torch.log(a / b)).sum(-1).mean()
        kl.backward()
        return zz.grad

    ga, gb = ours(), eager()
    cos = float((ga.flatten().double() @ gb.flatten().double()) / (ga.double().norm() * gb.double().norm()))
    out = dict(shape=f'{P} x {N} x {N} fp32 ({mb:.0f} MB per volume)', chain_ms_gd3=round(timed(ours), 3),
               chain_ms_eager_torch=round(timed(eager), 3), grad_cosine=round(cos, 7))
    _lib.profile_enable(True); _lib.profile_read()
    for _ in range(5): ours()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    # algorithmic bytes per launch: kl_map reads 2 volumes and writes 1 gradient; masked_cost_fwd reads 1, writes 1;
    # masked_cost_bwd reads grad_out and out, writes 1
    alg = dict(kl_map=3 * mb, masked_cost_fwd=2 * mb, masked_cost_bwd=3 * mb)
    out['kernels'] = {k: dict(us_per_launch=round(ms / cnt * 1e3, 1), gbs=round(alg[k] / (ms / cnt), 1) if k in alg else None)
                      for k, (cnt, ms) in prof.items()}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
