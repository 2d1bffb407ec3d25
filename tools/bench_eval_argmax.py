"""Keypoint -> arg-max pixel transfer at the reference's evaluation size (img 640 would need 1.26 GB per image on the
CPU oracle; 448 keeps the oracle within seconds).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '3d-vlm-gd_b200'))
import torch
from gd3 import _lib
from oracle import evaluate as oe

out = {}
for img, C, K in ((448, 768, 30), (840, 768, 30)):
    patch = stride = 14
    ph = 1 + (img - patch) // stride
    g = torch.Generator().manual_seed(img)
    d2 = torch.nn.functional.avg_pool2d(torch.randn(1, C, ph, ph, generator=g), 3, stride=1, padding=1)
    kd = torch.nn.functional.normalize(torch.randn(1, C, K, generator=g), dim=1)
    kdc, d2c = kd.cuda(), d2.cuda()
    for _ in range(3): idx = _lib.semantic_argmax(kdc, d2c, img, patch, stride)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): idx = _lib.semantic_argmax(kdc, d2c, img, patch, stride)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    rec = dict(img=img, C=C, K=K, patches=ph * ph, gpu_ms=round(ms, 4),
               reference_flops=2.0 * K * img * img * C, reference_map_bytes=4.0 * C * img * img)
    if img <= 448:
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        ref, sim = oe.semantic_argmax(kd, d2, img, patch, stride)
        rec['cpu_oracle_ms'] = round((time.perf_counter() - t0) * 1e3, 1)
        rec['cpu_threads'] = torch.get_num_threads()
        at = sim[torch.arange(K), idx.cpu()]
        rec['agree'] = int((idx.cpu() == ref).sum())
        rec['max_rel_gap'] = float(((sim.max(1).values - at).abs() / sim.max(1).values.abs().clamp_min(1)).max())
    out[f'img{img}'] = rec
print(json.dumps(out))
