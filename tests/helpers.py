"""Shared helpers for the parity tests."""
import numpy as np
import torch


def rel_err(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-12)


def cosine(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().flatten().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().flatten().cpu()
    na, nb = a.norm(), b.norm()
    if na == 0 and nb == 0:
        return 1.0
    if na == 0 or nb == 0:
        return 0.0
    return float((a @ b) / (na * nb))


def assert_grad_close(got, want, cos_min=0.999, name='grad', norm_rtol=2e-2):
    """Gradient parity bar of BASELINE.json: cosine >= 0.999 (plus a loose norm check)."""
    got = got.detach().float().cpu() if torch.is_tensor(got) else torch.as_tensor(got)
    want = want.detach().float().cpu() if torch.is_tensor(want) else torch.as_tensor(want)
    assert got.shape == want.shape, f'{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}'
    wn = float(want.double().norm())
    gn = float(got.double().norm())
    if wn < 1e-30:
        assert gn < 1e-12, f'{name}: expected zero gradient, got norm {gn}'
        return
    c = cosine(got, want)
    assert c >= cos_min, f'{name}: cosine {c} < {cos_min}'
    assert abs(gn - wn) <= norm_rtol * wn, f'{name}: norm {gn} vs {wn}'
