"""GPU parity: tcgen05 GEMM plumbing and the reciprocal-NN matcher (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import fast_nn as oracle_nn
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gd3mod():
    import gd3
    from gd3 import _lib
    _lib.load()
    return gd3


@pytest.mark.parametrize('tile_n', [128, 256, -128, -256])   # negative = 2-CTA (cta_group::2) kernel
@pytest.mark.parametrize('shape', [(1, 128, 256, 64), (2, 256, 512, 768), (3, 200, 333, 136), (1, 1369, 1369, 1024)])
def test_tc_gemm_matches_torch(gd3mod, tile_n, shape):
    """The hand-written tcgen05/TMA GEMM against a plain fp32 matmul of the same bf16 inputs."""
    from gd3 import _lib
    b, M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(b, M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(b, N, K, generator=g).to(torch.bfloat16).cuda()
    C = _lib.debug_gemm_bf16(A, B, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = torch.matmul(A.float().cpu(), B.float().cpu().transpose(1, 2))
    err = (C.cpu() - ref).abs().max().item()
    assert err <= 1e-3 * (K ** 0.5), f'max abs err {err}'


@pytest.mark.parametrize('a_mn,b_mn', [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize('tile_n', [128, 192, 256])
@pytest.mark.parametrize('shape', [(1, 128, 256, 64), (2, 256, 512, 768), (3, 200, 336, 136), (1, 1376, 1024, 1369)])
def test_tc_gemm_mn_major_operands(gd3mod, a_mn, b_mn, tile_n, shape):
    """MN-major UMMA descriptors (operands used without a transposed copy) against a plain fp32 matmul."""
    from gd3 import _lib
    b, M, N, K = shape
    g = torch.Generator().manual_seed(M + 3 * N + 5 * K)
    A = torch.randn(b, M, K, generator=g).to(torch.bfloat16)
    B = torch.randn(b, N, K, generator=g).to(torch.bfloat16)
    Ain = A.transpose(1, 2).contiguous() if a_mn else A
    Bin = B.transpose(1, 2).contiguous() if b_mn else B
    if (not (a_mn and b_mn)) and K % 8:
        pytest.skip('K-major operand needs K % 8 == 0')
    C = _lib.debug_gemm_bf16_mn(Ain.cuda(), Bin.cuda(), a_mn=a_mn, b_mn=b_mn, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = torch.matmul(A.float(), B.float().transpose(1, 2))
    err = (C.cpu() - ref).abs().max().item()
    assert err <= 1e-3 * (K ** 0.5), f'max abs err {err}'


def test_tc_gemm_exact_integers(gd3mod):
    """Small-integer operands: every partial sum is exact, so the result must be bit-identical."""
    from gd3 import _lib
    g = torch.Generator().manual_seed(5)
    A = torch.randint(-4, 5, (2, 384, 320), generator=g).float()
    B = torch.randint(-4, 5, (2, 272, 320), generator=g).float()
    for tile_n in (256, -256):
        C = _lib.debug_gemm_bf16(A.to(torch.bfloat16).cuda(), B.to(torch.bfloat16).cuda(), tile_n=tile_n)
        assert torch.equal(C.cpu(), A @ B.transpose(1, 2)), tile_n


@pytest.mark.parametrize('dist', ['dot', 'l2'])
def test_reciprocal_nn_golden(gd3mod, golden, dist):
    from gd3.compat import fast_nn
    g = golden('fast_nn.npz')
    A, B = g['exact/A'], g['exact/B']
    for blk in (None, 128 if dist == 'dot' else 100):
        a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist=dist, block_size=blk)
        assert a.dtype == np.int64 and b.dtype == np.int64
        assert (a == g[f'exact/{dist}/nnA']).all()
        assert (b == g[f'exact/{dist}/nnB']).all()


def test_reciprocal_nn_cfg3_exact_set(gd3mod):
    """BASELINE.json config 3: 8192 x 8192 x 24, bit-exact indices on the exactly-representable set."""
    from gd3.compat import fast_nn
    A = synth.nn_exact_set(301, 8192)
    B = synth.nn_exact_set(302, 8192)
    a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist='dot', block_size=2 ** 13)
    ra, rb = oracle_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot', block_size=2 ** 13)
    assert (a == ra).all() and (b == rb).all()
    # size-independent property: the reported neighbour attains the row / column maximum
    S = A @ B.T
    assert torch.equal(S[torch.arange(8192), torch.from_numpy(a)], S.max(dim=1).values)
    assert torch.equal(S[torch.from_numpy(b), torch.arange(8192)], S.max(dim=0).values)


def test_reciprocal_nn_real_set_near_ties_only(gd3mod):
    """Gaussian unit descriptors: any index mismatch vs the CPU matmul must be a rounding-level near-tie."""
    from gd3.compat import fast_nn
    A = synth.nn_real_set(303, 4096)
    B = synth.nn_real_set(304, 5000)
    a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist='dot')
    ra, rb = oracle_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot')
    S = (A.double() @ B.double().T)
    bad = np.nonzero(a != ra)[0]
    for i in bad:
        assert abs(S[i, a[i]] - S[i, ra[i]]) < 1e-6
    bad_b = np.nonzero(b != rb)[0]
    for j in bad_b:
        assert abs(S[b[j], j] - S[rb[j], j]) < 1e-6
    assert len(bad) <= 4 and len(bad_b) <= 4


def test_reciprocal_nn_edge_cases(gd3mod):
    from gd3 import _lib
    from gd3.compat import fast_nn
    A = synth.nn_exact_set(1, 37)
    B = synth.nn_exact_set(2, 5)
    with pytest.raises(ValueError):
        fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist='cosine')
    a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist='dot')
    ra, rb = oracle_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot')
    assert (a == ra).all() and (b == rb).all()
    # duplicated rows -> ties -> lowest index
    D = torch.cat([B, B, B])
    a, _ = fast_nn.bruteforce_reciprocal_nns(B, D, device='cuda', dist='l2')
    assert (a == np.arange(5)).all()
    # empty query (cdistMatcher contract, mast3r/fast_nn.py:80-81)
    m = fast_nn.cdistMatcher(B, device='cuda')
    assert m.query(torch.empty(0, 24)) == (None, [])
    with pytest.raises(_lib.Gd3Error):
        fast_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot')


def test_fast_reciprocal_nns_golden(gd3mod, golden):
    from gd3.compat import fast_nn
    g = golden('fast_nn.npz')
    d1, d2 = torch.from_numpy(g['maps/d1']), torch.from_numpy(g['maps/d2'])
    for tag, kw in (('s8', dict(subsample_or_initxy1=8)), ('s4', dict(subsample_or_initxy1=4))):
        xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, device='cuda', dist='dot', block_size=2 ** 10, **kw)
        assert (xy1 == g[f'maps/{tag}/xy1']).all() and (xy2 == g[f'maps/{tag}/xy2']).all()
    i1, i2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=8, ret_xy=False, device='cuda', dist='dot')
    assert (i1 == g['maps/s8_idx/i1']).all() and (i2 == g['maps/s8_idx/i2']).all()
    xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=(g['maps/seeds/x'], g['maps/seeds/y']),
                                           pixel_tol=3, device='cuda', dist='dot')
    assert (xy1 == g['maps/seeds_tol3/xy1']).all() and (xy2 == g['maps/seeds_tol3/xy2']).all()
    xy1, xy2, basin = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=8, ret_basin=True, device='cuda',
                                                  dist='dot')
    assert (xy1 == g['maps/basin/xy1']).all() and (basin == g['maps/basin/basin']).all()


def test_fast_reciprocal_nns_real_shape(gd3mod):
    """MASt3R-sized maps (384 x 512 x 24, S=16) against the CPU oracle on exactly-representable descriptors."""
    from gd3.compat import fast_nn
    d1, d2 = synth.nn_desc_maps(305, 384, 512)
    d1 = torch.round(d1 * 16) / 8
    d2 = torch.round(d2 * 16) / 8
    xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=16, device='cuda', dist='dot',
                                           block_size=2 ** 13)
    r1, r2 = oracle_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=16, device='cpu', dist='dot',
                                           block_size=2 ** 13)
    assert xy1.shape == r1.shape and (xy1 == r1).all() and (xy2 == r2).all()


@pytest.mark.parametrize('dist', ['dot', 'l2'])
@pytest.mark.parametrize('nq,dim', [(1, 24), (7, 24), (16, 24), (17, 24), (64, 24), (5, 3), (33, 10), (12, 128)])
def test_few_queries_streaming_kernel(gd3mod, dist, nq, dim):
    """A handful of queries against a large DB takes the streaming kernel: same answers (lowest-index ties) as the
    oracle, and as the tile kernel on the same data (bit-identical scores)."""
    from gd3 import _lib
    g = torch.Generator().manual_seed(nq * 131 + dim)
    DB = torch.randint(-8, 9, (6000, dim), generator=g).float() / 8          # exactly representable, many ties
    DB[4000:4100] = DB[100:200]                                               # duplicated rows
    Q = DB[torch.randint(0, 6000, (nq,), generator=g)] + (torch.randint(-1, 2, (nq, dim), generator=g).float() / 8)
    ref, _ = oracle_nn.bruteforce_reciprocal_nns(Q, DB, device='cpu', dist=dist)
    got, _ = _lib.reciprocal_nn(Q.cuda(), DB.cuda(), dist=dist, want_B=False)        # streaming kernel (nq <= 64)
    assert (got.cpu().numpy() == ref).all()
    both, _ = _lib.reciprocal_nn(Q.cuda(), DB.cuda(), dist=dist, want_B=True)        # tile kernel
    assert torch.equal(both, got)


@pytest.mark.parametrize('dist', ['dot', 'l2'])
def test_device_resident_ping_pong(gd3mod, dist):
    """gd3_fast_reciprocal_nn against the reference loop (oracle) for both distances, grid and explicit seeds."""
    from gd3.compat import fast_nn
    d1, d2 = synth.nn_desc_maps(77, 96, 128)
    d1 = torch.round(d1 * 16) / 8
    d2 = torch.round(d2 * 16) / 8
    for S in (16, 8, 3):
        xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=S, device='cuda', dist=dist)
        r1, r2 = oracle_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=S, device='cpu', dist=dist, block_size=2 ** 13)
        assert xy1.shape == r1.shape and (xy1 == r1).all() and (xy2 == r2).all(), S
    # explicit seeds: a single round (mast3r/fast_nn.py:122-128)
    ys, xs = np.mgrid[2:96:7, 3:128:5].reshape(2, -1)
    xy1, xy2 = fast_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=(xs, ys), device='cuda', dist=dist)
    r1, r2 = oracle_nn.fast_reciprocal_NNs(d1, d2, subsample_or_initxy1=(xs, ys), device='cpu', dist=dist,
                                           block_size=2 ** 13)
    assert xy1.shape == r1.shape and (xy1 == r1).all() and (xy2 == r2).all()


def test_device_ping_pong_poll_modes_agree(gd3mod):
    """host_poll only decides when to stop launching rounds: both modes return identical arrays."""
    from gd3 import _lib
    d1, d2 = synth.nn_desc_maps(78, 64, 96)
    p1 = (torch.round(d1 * 16) / 8).reshape(-1, 24).cuda()
    p2 = (torch.round(d2 * 16) / 8).reshape(-1, 24).cuda()
    seeds = torch.arange(5, 64 * 96, 37, dtype=torch.int32).cuda()
    a = _lib.fast_reciprocal_nn(p1, p2, seeds, max_iter=10, dist='dot', host_poll=True)
    b = _lib.fast_reciprocal_nn(p1, p2, seeds, max_iter=10, dist='dot', host_poll=False)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert a[2].any()


def test_extract_correspondences_nonsym_golden(gd3mod, golden):
    """Both-sided matching with confidences (mast3r/fast_nn.py:191-223) against the live reference's output."""
    from gd3.compat import fast_nn
    g = golden('fast_nn_extra.npz')
    d1, d2 = torch.from_numpy(g['d1']), torch.from_numpy(g['d2'])
    for tag, tol in (('tol0', 0), ('tol2', 2)):
        xy1, xy2, conf = fast_nn.extract_correspondences_nonsym(d1, d2, torch.from_numpy(g['cA']), g['cB'], subsample=8,
                                                                device='cuda', pixel_tol=tol)
        assert xy1.is_cuda and conf.is_cuda
        assert (xy1.cpu().numpy() == g[f'{tag}/xy1']).all() and (xy2.cpu().numpy() == g[f'{tag}/xy2']).all()
        assert np.array_equal(conf.cpu().numpy(), g[f'{tag}/conf'])


@pytest.mark.parametrize('seed', range(10))
def test_reciprocal_nn_random_sizes(gd3mod, seed):
    """Random sizes / descriptor widths (incl. 1, odd and > 128) on exactly-representable descriptors: indices equal
    the oracle's in both directions and for both distances (tools/probe_nn_shapes.py runs longer sweeps)."""
    import random
    from gd3.compat import fast_nn
    rnd = random.Random(500 + seed)
    na = rnd.choice([1, 2, 31, 64, 65, 127, 129, 500, 1000, 2049, 3000])
    nb = rnd.choice([1, 3, 33, 128, 130, 777, 1024, 2500, 4097])
    dim = rnd.choice([1, 2, 3, 5, 8, 24, 25, 64, 100, 128, 130, 256])
    dist = rnd.choice(['dot', 'l2'])
    A = synth.nn_exact_set(10 * seed + 1, na, dim=dim, dup=min(8, na // 2))
    B = synth.nn_exact_set(10 * seed + 2, nb, dim=dim, dup=min(8, nb // 2))
    a, b = fast_nn.bruteforce_reciprocal_nns(A, B, device='cuda', dist=dist)
    ra, rb = oracle_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist=dist)
    assert (a == ra).all() and (b == rb).all(), (na, nb, dim, dist)


@pytest.mark.parametrize('shift_a,shift_b', [(1, 0), (0, 1), (3, 2)])
def test_reciprocal_nn_misaligned_rows(gd3mod, shift_a, shift_b):
    """24-d descriptors whose base pointers are not 16-byte aligned (views into a larger buffer at an odd float offset):
    the tile kernel's loaders fall back from 16-byte to 4-byte copies; indices stay identical to the aligned call."""
    from gd3 import _lib
    A = synth.nn_exact_set(71, 700)
    B = synth.nn_exact_set(72, 900)
    ra, rb = oracle_nn.bruteforce_reciprocal_nns(A, B, device='cpu', dist='dot')

    def shifted(x, s):
        buf = torch.zeros(x.numel() + 8, device='cuda')
        v = buf[s:s + x.numel()].view(x.shape)
        v.copy_(x)
        assert v.data_ptr() % 16 == (4 * s) % 16
        return v

    a, b = _lib.reciprocal_nn(shifted(A, shift_a), shifted(B, shift_b), dist='dot')
    assert (a.cpu().numpy() == ra).all() and (b.cpu().numpy() == rb).all()
