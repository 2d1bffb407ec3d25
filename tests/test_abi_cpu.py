"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol include/gd3.h declares,
the ctypes table matches the header, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'gd3.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    decls = re.findall(r'\b([A-Za-z_][A-Za-z0-9_ \*]*?)\b(gd3_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;', src)
    return {name: args for _, name, args in decls}


def test_library_builds_and_exports_every_header_symbol():
    import __graft_entry__
    lib_path = __graft_entry__.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    fns = header_functions()
    assert len(fns) >= 14
    for name in fns:
        assert hasattr(lib, name), f'{name} declared in include/gd3.h but not exported'


def test_ctypes_table_matches_header():
    from gd3 import _lib
    fns = header_functions()
    assert set(_lib.SYMBOLS) == set(fns), set(_lib.SYMBOLS) ^ set(fns)
    for name, args in fns.items():
        n_args = 0 if args.strip() in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        assert len(_lib.SYMBOLS[name][1]) == n_args, f'{name}: header has {n_args} parameters'


def test_version_and_error_channel():
    from gd3 import _lib
    lib = _lib.load()
    assert lib.gd3_version() == 100
    assert isinstance(lib.gd3_last_error(), bytes)
    assert lib.gd3_launch_count() >= 0
    # pure host-side size queries work without a device
    assert lib.gd3_cost_kl_workspace(32, 1024, 768, 0, 1) > 0
    assert lib.gd3_reciprocal_nn_workspace(8192, 8192) >= 2 * 8192 * 8
    assert lib.gd3_smooth_ap_workspace(4, 512, 768, 1) > lib.gd3_smooth_ap_workspace(4, 512, 768, 0)
    assert lib.gd3_depth_head_loss_workspace(8, 300, 1024, 1, 1) > 0
    assert lib.gd3_kl_divergence_map_workspace(32 * 1024) >= 32 * 1024 * 4
    assert lib.gd3_cost_kl_group_size(32, 1024, 768, 5) == 5
    assert 1 <= lib.gd3_cost_kl_group_size(32, 1024, 768, 0) <= 32


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly (never route through the oracle)."""
    from gd3 import _lib, ops
    from gd3.compat import fast_nn, functions, losses
    f = torch.randn(1, 16, 8)
    t = torch.softmax(torch.randn(1, 16, 16), -1)
    with pytest.raises(_lib.Gd3Error):
        ops.cost_volume_kl(f, f, t, t)
    with pytest.raises(_lib.Gd3Error):
        ops.smooth_ap(f, f, torch.zeros(1, 16, 3), torch.zeros(1, 16, 3))
    with pytest.raises(_lib.Gd3Error):
        ops.sample_tokens(f, (4, 4), torch.zeros(1, 3, 2))
    with pytest.raises(_lib.Gd3Error):
        fast_nn.bruteforce_reciprocal_nns(torch.randn(4, 3), torch.randn(5, 3), device='cuda', dist='dot')
    with pytest.raises(_lib.Gd3Error):
        fast_nn.bruteforce_reciprocal_nns(torch.randn(4, 3), torch.randn(5, 3), device='cpu', dist='dot')
    with pytest.raises(_lib.Gd3Error):
        functions.interpolate_features(torch.randn(1, 8, 4, 4), torch.zeros(1, 3, 2), 56, 56)
    from oracle.losses import DepthHead
    with pytest.raises(_lib.Gd3Error):
        losses.pairwise_logistic_ranking_loss(DepthHead(8), f, torch.rand(1, 16))
    with pytest.raises(_lib.Gd3Error):
        losses.kl_divergence_map(t, t)
    with pytest.raises(_lib.Gd3Error):
        functions.get_masked_patch_cost(t, torch.ones(16, dtype=torch.bool))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, '3d-vlm-gd_b200', 'gd3')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
                assert 'bench_common' not in src, fn
