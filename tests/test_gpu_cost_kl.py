"""GPU parity of the fused cost-volume KL loss (through the C ABI) against the golden vectors of the
live reference and against the CPU oracle.  Bars (BASELINE.json): loss rel. err <= 1e-3, gradient
cosine >= 0.999."""
import pytest
import torch

from oracle import bodies, synth
from helpers import assert_grad_close, rel_err
from test_oracle_golden import KL_CASES, kl_inputs

pytestmark = pytest.mark.gpu
T = torch.as_tensor
LOSS_RTOL = 1e-3


@pytest.fixture(scope='module')
def ops():
    from gd3 import ops as o
    return o


def run_gpu(ops, f1, f2, t12, t21, m1, m2, variant, dtype=torch.float32, **kw):
    d = 'cuda'
    F1 = f1.to(d, dtype).requires_grad_(True)
    F2 = f2.to(d, dtype).requires_grad_(True)
    loss = ops.cost_volume_kl(F1, F2, t12.to(d), t21.to(d), m1.to(d), m2.to(d), variant=variant, **kw)
    loss.sum().backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), F1.grad.float().cpu(), F2.grad.float().cpu()


@pytest.mark.parametrize('case', KL_CASES)
def test_cost_kl_golden(ops, golden, case):
    g = golden('cost_kl.npz')
    f1, f2, t12, t21, m1, m2, variant = kl_inputs(g[f'{case}/meta'])
    loss, g1, g2 = run_gpu(ops, f1[None], f2[None], t12[None], t21[None], m1[None], m2[None], variant)
    want = float(g[f'{case}/loss'])
    assert abs(float(loss[0]) - want) <= LOSS_RTOL * abs(want) + 1e-7, (float(loss[0]), want)
    assert_grad_close(g1[0], T(g[f'{case}/g1']), name='g1')
    assert_grad_close(g2[0], T(g[f'{case}/g2']), name='g2')


def oracle_pair(f1, f2, t12, t21, m1, m2, variant):
    a = f1.clone().requires_grad_(True)
    b = f2.clone().requires_grad_(True)
    loss = bodies.cost_volume_kl(a, b, t12, t21, m1, m2, variant)
    ga, gb = torch.autograd.grad(loss, [a, b])
    return float(loss), ga, gb


def test_cost_kl_batched_groups_and_dtypes(ops):
    """P = 5 pairs with different masks; several group sizes (incl. a ragged last group), fp32 and bf16 I/O."""
    N, C, P = 200, 96, 5
    pairs = []
    for p in range(P):
        f1, f2 = synth.features(900 + p, N, C)
        mode = ['bernoulli', 'all', 'bernoulli', 'none', 'bernoulli'][p]
        pairs.append((f1, f2, synth.teacher_volume(910 + p, N, 'mast3r'), synth.teacher_volume(920 + p, N, 'mast3r'),
                      synth.patch_mask(930 + p, N, mode=mode), synth.patch_mask(940 + p, N, mode=mode)))
    stack = [torch.stack([q[k] for q in pairs]) for k in range(6)]
    want = [oracle_pair(*q, 'mast3r') for q in pairs]
    for dtype in (torch.float32, torch.bfloat16):
        for ppg in (0, 1, 2, 5):
            loss, g1, g2 = run_gpu(ops, *stack, 'mast3r', dtype=dtype, pairs_per_group=ppg)
            for p in range(P):
                assert abs(float(loss[p]) - want[p][0]) <= LOSS_RTOL * abs(want[p][0]) + 1e-7, (dtype, ppg, p)
                assert_grad_close(g1[p], want[p][1], name=f'g1[{p}]', norm_rtol=3e-2)
                assert_grad_close(g2[p], want[p][2], name=f'g2[{p}]', norm_rtol=3e-2)


def test_cost_kl_channel_major_view_and_forward_only(ops):
    """The MASt3R path hands over a (1, N, C) view with N-stride 1 / C-stride N."""
    N, C = 144, 80
    f1, f2 = synth.features(77, N, C)
    t12, t21 = synth.teacher_volume(78, N, 'vggt'), synth.teacher_volume(79, N, 'vggt')
    m1, m2 = synth.patch_mask(80, N), synth.patch_mask(81, N)
    want, ga, gb = oracle_pair(f1, f2, t12, t21, m1, m2, 'vggt')
    F1 = f1.t().contiguous().cuda().t()[None].requires_grad_(True)      # strides (.., 1, N)
    F2 = f2.cuda()[None].requires_grad_(True)
    assert F1.stride(1) == 1 and F1.stride(2) == N
    loss = ops.cost_volume_kl(F1, F2, t12.cuda()[None], t21.cuda()[None], m1.cuda(), m2.cuda(), variant='vggt')
    loss.sum().backward()
    assert rel_err(loss[0].item(), want) <= LOSS_RTOL
    assert_grad_close(F1.grad[0], ga, name='g1')
    assert_grad_close(F2.grad[0], gb, name='g2')
    with torch.no_grad():
        l2 = ops.cost_volume_kl(F1, F2, t12.cuda()[None], t21.cuda()[None], m1.cuda(), m2.cuda(), variant='vggt')
    assert rel_err(l2[0].item(), want) <= LOSS_RTOL
    # upstream gradient scaling
    F2.grad = None
    (3.0 * ops.cost_volume_kl(F1.detach(), F2, t12.cuda()[None], t21.cuda()[None], m1.cuda(), m2.cuda(),
                              variant='vggt')).sum().backward()
    assert_grad_close(F2.grad[0], 3.0 * gb, name='scaled g2')


@pytest.mark.parametrize('cfg', ['cfg2', 'cfg4'])
def test_cost_kl_full_size_pair(ops, cfg):
    """One pair at the BASELINE.json sizes (N=1024,C=768 / N=1369,C=1024) against the CPU oracle."""
    c = synth.CONFIGS[cfg]
    N, C, variant = c['N'], c['C'], c['variant']
    f1, f2 = synth.features(5000, N, C)
    t12, t21 = synth.teacher_volume(5001, N, variant, heads=4), synth.teacher_volume(5002, N, variant, heads=4)
    m1, m2 = synth.patch_mask(5003, N), synth.patch_mask(5004, N)
    want, ga, gb = oracle_pair(f1, f2, t12, t21, m1, m2, variant)
    loss, g1, g2 = run_gpu(ops, f1[None], f2[None], t12[None], t21[None], m1[None], m2[None], variant,
                           dtype=torch.bfloat16)
    assert rel_err(loss[0].item(), want) <= LOSS_RTOL, (loss[0].item(), want)
    assert_grad_close(g1[0], ga, name='g1', norm_rtol=3e-2)
    assert_grad_close(g2[0], gb, name='g2', norm_rtol=3e-2)


def test_cost_kl_properties_at_batch_size(ops):
    """Size-independent checks on a 6-pair cfg2-shaped batch: identical pairs give identical results
    whatever their position / group, swapping the views swaps the gradients, and rows of the
    gradient are orthogonal to the features (the loss only sees normalised features)."""
    c = synth.CONFIGS['cfg2']
    N, C = c['N'], c['C']
    f1, f2 = synth.features(6000, N, C)
    t12, t21 = synth.teacher_volume(6001, N, 'mast3r'), synth.teacher_volume(6002, N, 'mast3r')
    m1, m2 = synth.patch_mask(6003, N), synth.patch_mask(6004, N)
    P = 6
    rep = lambda x: x[None].expand(P, *x.shape).contiguous()
    F1, F2, T12, T21, M1, M2 = map(rep, (f1, f2, t12, t21, m1, m2))
    # pair 3 is the same problem with the two views exchanged
    F1[3], F2[3], T12[3], T21[3], M1[3], M2[3] = f2, f1, t21, t12, m2, m1
    loss, g1, g2 = run_gpu(ops, F1, F2, T12, T21, M1, M2, 'mast3r', dtype=torch.bfloat16, pairs_per_group=4)
    for p in (1, 2, 4, 5):
        assert rel_err(loss[p].item(), loss[0].item()) < 1e-5
        assert_grad_close(g1[p], g1[0], cos_min=0.99999, name='repeat', norm_rtol=1e-3)
    assert rel_err(loss[3].item(), loss[0].item()) < 1e-4
    assert_grad_close(g1[3], g2[0], cos_min=0.9999, name='swap', norm_rtol=1e-2)
    assert_grad_close(g2[3], g1[0], cos_min=0.9999, name='swap', norm_rtol=1e-2)
    radial = (g1[0] * f1).sum(-1).abs().max() / (g1[0].norm(dim=-1).max() * f1.norm(dim=-1).max())
    assert radial < 2e-2


@pytest.mark.parametrize('cfg', ['cfg2', 'cfg4'])
@pytest.mark.parametrize('with_stats', [True, False])
def test_cost_kl_packed_teacher_full_size(ops, cfg, with_stats):
    """Teacher volumes in the producers' packed form (fp16 * 1024 + row statistics, gd3_teacher_pack) against the CPU
    oracle on the ORIGINAL fp32 volumes at the BASELINE.json sizes: same bars as the fp32 path."""
    c = synth.CONFIGS[cfg]
    N, C, variant = c['N'], c['C'], c['variant']
    f1, f2 = synth.features(5100, N, C)
    t12, t21 = synth.teacher_volume(5101, N, variant, heads=4), synth.teacher_volume(5102, N, variant, heads=4)
    m1, m2 = synth.patch_mask(5103, N), synth.patch_mask(5104, N)
    want, ga, gb = oracle_pair(f1, f2, t12, t21, m1, m2, variant)
    p12, s12 = ops.pack_teacher(t12.cuda()[None])
    p21, s21 = ops.pack_teacher(t21.cuda()[None])
    assert p12.dtype == torch.float16 and s12.shape == (1, 3, N)
    # the statistics are those of the fp32 rows
    R = t12.double().sum(-1)
    tt = (t12.double() / R.clamp_min(1e-8)[:, None]).clamp_min(1e-8)
    assert torch.allclose(s12[0, 0].double().cpu(), R, rtol=1e-5)
    assert torch.allclose(s12[0, 1].double().cpu(), tt.sum(-1), rtol=1e-5)
    assert torch.allclose(s12[0, 2].double().cpu(), (tt * tt.log()).sum(-1), rtol=1e-4, atol=1e-6)
    F1 = f1.cuda().to(torch.bfloat16)[None].requires_grad_(True)
    F2 = f2.cuda().to(torch.bfloat16)[None].requires_grad_(True)
    a12, a21 = ((p12, s12), (p21, s21)) if with_stats else (p12, p21)
    loss = ops.cost_volume_kl(F1, F2, a12, a21, m1.cuda()[None], m2.cuda()[None], variant=variant)
    loss.sum().backward()
    assert rel_err(loss[0].item(), want) <= LOSS_RTOL, (loss[0].item(), want)
    assert_grad_close(F1.grad[0].float().cpu(), ga, name='g1', norm_rtol=3e-2)
    assert_grad_close(F2.grad[0].float().cpu(), gb, name='g2', norm_rtol=3e-2)
    with torch.no_grad():
        fwd = ops.cost_volume_kl(F1, F2, a12, a21, m1.cuda()[None], m2.cuda()[None], variant=variant)
    assert rel_err(fwd[0].item(), want) <= LOSS_RTOL


def test_cost_kl_packed_teacher_ragged_batched(ops):
    """Packed teachers on a ragged size (N = 15 * 13, rows not 8-byte aligned -> scalar loads), 3 pairs, both variants,
    all-masked and all-kept rows; mixing forms is rejected."""
    N, C, P = 195, 72, 3
    for variant in ('mast3r', 'vggt'):
        pairs = []
        for p in range(P):
            f1, f2 = synth.features(5200 + p, N, C)
            mode = ['bernoulli', 'all', 'none'][p]
            pairs.append((f1, f2, synth.teacher_volume(5210 + p, N, variant), synth.teacher_volume(5220 + p, N, variant),
                          synth.patch_mask(5230 + p, N, mode=mode), synth.patch_mask(5240 + p, N, mode=mode)))
        want = [oracle_pair(*q, variant) for q in pairs]
        st = [torch.stack([q[k] for q in pairs]).cuda() for k in range(6)]
        t12 = ops.pack_teacher(st[2])
        t21 = ops.pack_teacher(st[3])
        F1, F2 = st[0].requires_grad_(True), st[1].requires_grad_(True)
        loss = ops.cost_volume_kl(F1, F2, t12, t21, st[4], st[5], variant=variant, pairs_per_group=2)
        loss.sum().backward()
        for p in range(P):
            assert abs(float(loss[p]) - want[p][0]) <= LOSS_RTOL * abs(want[p][0]) + 1e-7, (variant, p, float(loss[p]), want[p][0])
            if float(want[p][1].norm()) > 0:
                assert_grad_close(F1.grad[p].cpu(), want[p][1], name=f'g1[{p}]', norm_rtol=3e-2)
                assert_grad_close(F2.grad[p].cpu(), want[p][2], name=f'g2[{p}]', norm_rtol=3e-2)
    with pytest.raises(ValueError):
        ops.cost_volume_kl(F1, F2, t12, st[3], st[4], st[5], variant='vggt')


@pytest.mark.parametrize('N,C', [(130, 72), (257, 200), (64, 8), (333, 136)])
def test_cost_kl_tma_store_partial_chunks(ops, N, C):
    """bf16 gradients leave through TMA stores in 32-column boxes: channel counts that end inside a box (C = 72, 200, 8,
    136), token counts that end inside a 32-row box, packed and fp32 teachers, against the CPU oracle."""
    P = 2
    pairs = []
    for p in range(P):
        f1, f2 = synth.features(5300 + p + N, N, C)
        pairs.append((f1, f2, synth.teacher_volume(5310 + p, N, 'vggt'), synth.teacher_volume(5320 + p, N, 'vggt'),
                      synth.patch_mask(5330 + p, N), synth.patch_mask(5340 + p, N)))
    want = [oracle_pair(*q, 'vggt') for q in pairs]
    st = [torch.stack([q[k] for q in pairs]).cuda() for k in range(6)]
    for packed in (False, True):
        F1 = st[0].to(torch.bfloat16).requires_grad_(True)
        F2 = st[1].to(torch.bfloat16).requires_grad_(True)
        t12, t21 = (ops.pack_teacher(st[2]), ops.pack_teacher(st[3])) if packed else (st[2], st[3])
        loss = ops.cost_volume_kl(F1, F2, t12, t21, st[4], st[5], variant='vggt')
        loss.sum().backward()
        for p in range(P):
            assert rel_err(loss[p].item(), want[p][0]) <= LOSS_RTOL, (N, C, packed, p, loss[p].item(), want[p][0])
            assert torch.isfinite(F1.grad[p].float()).all() and torch.isfinite(F2.grad[p].float()).all()
            assert_grad_close(F1.grad[p].float().cpu(), want[p][1], name=f'g1[{p}]', norm_rtol=3e-2)
            assert_grad_close(F2.grad[p].float().cpu(), want[p][2], name=f'g2[{p}]', norm_rtol=3e-2)
