"""ctypes binding of lib3dgd.so (the C ABI declared in include/gd3.h).

This file is the reference-side stub INTEGRATION.md describes: plain pointers and sizes in, error
codes out.  Torch is used only to obtain device pointers, the current stream and workspace memory.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GD3_LIB', os.path.join(os.path.dirname(_HERE), 'lib', 'lib3dgd.so'))

# every symbol include/gd3.h declares: name -> (restype, argtypes)
_c = ctypes
_i64, _int, _vp, _sz, _f32 = _c.c_int64, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_float
SYMBOLS = {
    'gd3_version': (_int, []),
    'gd3_last_error': (_c.c_char_p, []),
    'gd3_launch_count': (_c.c_longlong, []),
    'gd3_profile_enable': (None, [_int]),
    'gd3_profile_read': (_sz, [_c.c_char_p, _sz]),
    'gd3_reciprocal_nn_workspace': (_sz, [_i64, _i64]),
    'gd3_reciprocal_nn': (_int, [_vp, _i64, _vp, _i64, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    'gd3_kp_prepare': (_int, [_vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _i64, _vp, _vp, _vp]),
    'gd3_masked_patch_cost': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _int, _f32, _f32, _vp, _vp, _vp]),
    'gd3_masked_patch_cost_backward': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _int, _f32, _f32, _vp, _vp]),
    'gd3_kl_divergence_map_workspace': (_sz, [_i64]),
    'gd3_kl_divergence_map': (_int, [_vp, _vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd3_point_cloud_to_depth_workspace': (_sz, [_i64, _i64, _i64]),
    'gd3_point_cloud_to_depth': (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    'gd3_vggt_attn_workspace': (_sz, [_i64, _i64, _i64]),
    'gd3_vggt_attn_accumulate': (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _f32, _int, _f32, _int, _vp, _vp, _vp, _sz,
                                        _vp]),
    'gd3_teacher_volume_workspace': (_sz, [_i64, _i64, _i64]),
    'gd3_teacher_volume': (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f32, _int, _vp, _vp, _sz, _vp]),
    'gd3_semantic_argmax_workspace': (_sz, [_i64, _i64, _i64]),
    'gd3_semantic_argmax': (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    'gd3_fast_reciprocal_nn_workspace': (_sz, [_i64, _i64]),
    'gd3_fast_reciprocal_nn': (_int, [_vp, _i64, _vp, _i64, _i64, _int, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd3_cost_kl_group_size': (_i64, [_i64, _i64, _i64, _i64]),
    'gd3_cost_kl_workspace': (_sz, [_i64, _i64, _i64, _i64, _int]),
    'gd3_cost_kl': (_int, [_vp, _vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _int, _f32,
                           _vp, _vp, _i64, _i64, _vp, _vp, _int, _f32, _f32, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    'gd3_teacher_pack': (_int, [_vp, _i64, _i64, _i64, _i64, _f32, _f32, _vp, _i64, _vp, _vp]),
    'gd3_smooth_ap_workspace': (_sz, [_i64, _i64, _i64, _int]),
    'gd3_smooth_ap': (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp,
                             _sz, _vp]),
    'gd3_infonce_workspace': (_sz, [_i64, _i64, _i64, _int]),
    'gd3_infonce': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd3_depth_head_loss_workspace': (_sz, [_i64, _i64, _i64, _int, _int]),
    'gd3_depth_head_loss': (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _f32, _int,
                                   _f32, _f32, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gd3_sample_tokens_fwd': (_int, [_vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp,
                                     _i64, _int, _int, _int, _vp, _i64, _i64, _i64, _vp, _vp]),
    'gd3_sample_tokens_bwd': (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _i64,
                                     _i64, _i64, _i64, _i64, _int, _int, _int, _vp, _i64, _i64, _i64, _i64, _vp, _i64,
                                     _i64, _i64, _vp]),
    'gd3_debug_gemm_bf16': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int, _vp]),
    'gd3_debug_gemm_bf16_mn': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _int, _vp]),
}

DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
DIST = {'dot': 0, 'l2': 1}
VARIANT = {'mast3r': 0, 'vggt': 1, 'me': 2, 'me_joint': 3}

_lib = None


class Gd3Error(RuntimeError):
    pass


def load():
    """Load lib3dgd.so once.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Gd3Error(f'lib3dgd.so not found at {LIB_PATH}; build it with '
                           f'`python 3d-vlm-gd_b200/build.py` (there is no CPU / PyTorch fallback)')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().gd3_last_error().decode(errors='replace')
        if rc == -1:
            raise ValueError(msg)
        raise Gd3Error(f'lib3dgd error {rc}: {msg}')


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise Gd3Error('gd3 needs a CUDA device (sm_100a); there is no CPU fallback')
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError('gd3 ops take CUDA tensors')


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes, device):
    """Workspace from torch's caching allocator (caller-owned as the C ABI requires)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def dtype_code(t):
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise ValueError(f'unsupported feature dtype {t.dtype} (float32 or bfloat16)')


# ---------------------------------------------------------------------------------------------
def masked_patch_cost(cost, mask1, mask2, use_softmax, eps, temperature, want_row_sum=True):
    """out (B, hw, hw2) fp32 and the per-row sums (B * hw,) of ``get_masked_patch_cost`` for a CUDA volume."""
    require_cuda(cost, mask1, mask2)
    lib = load()
    cost = cost.contiguous().float()
    B, hw, hw2 = cost.shape
    m1 = mask1.to(torch.bool).contiguous()
    m2 = None if mask2 is None else mask2.to(torch.bool).contiguous()
    if m1.numel() != hw or (m2 is not None and m2.numel() != hw2):
        raise ValueError('get_masked_patch_cost: mask sizes do not match the volume')
    out = torch.empty_like(cost)
    row_sum = torch.empty(B * hw, dtype=torch.float32, device=cost.device) if want_row_sum else None
    with torch.cuda.device(cost.device):
        check(lib.gd3_masked_patch_cost(ptr(cost), B, hw, hw2, ptr(m1), ptr(m2), int(bool(use_softmax)), float(eps),
                                        float(temperature), ptr(out), ptr(row_sum), stream_ptr()))
    return out, row_sum, m1, m2


def masked_patch_cost_backward(grad_out, out, row_sum, m1, m2, use_softmax, eps, temperature):
    lib = load()
    grad_out = grad_out.contiguous().float()
    B, hw, hw2 = out.shape
    grad_cost = torch.empty_like(out)
    with torch.cuda.device(out.device):
        check(lib.gd3_masked_patch_cost_backward(ptr(grad_out), ptr(out), ptr(row_sum), B, hw, hw2, ptr(m1), ptr(m2),
                                                 int(bool(use_softmax)), float(eps), float(temperature), ptr(grad_cost),
                                                 stream_ptr()))
    return grad_cost


def kl_divergence_map(teacher, student, eps, want_grad_student, want_grad_teacher):
    """loss (0-d fp32) of ``kl_divergence_map`` for CUDA volumes of equal shape (..., n), and the two gradients (or None)."""
    require_cuda(teacher, student)
    if teacher.shape != student.shape:
        raise ValueError(f'kl_divergence_map: shapes differ ({tuple(teacher.shape)} vs {tuple(student.shape)})')
    lib = load()
    t = teacher.contiguous().float()
    s = student.contiguous().float()
    n = t.shape[-1]
    rows = t.numel() // max(n, 1)
    loss = torch.empty((), dtype=torch.float32, device=t.device)
    gs = torch.empty_like(s) if want_grad_student else None
    gt = torch.empty_like(t) if want_grad_teacher else None
    ws_bytes = lib.gd3_kl_divergence_map_workspace(rows)
    ws = workspace(ws_bytes, t.device)
    with torch.cuda.device(t.device):
        check(lib.gd3_kl_divergence_map(ptr(t), ptr(s), rows, n, float(eps), ptr(loss), ptr(gs), ptr(gt), ptr(ws),
                                        ws.numel(), stream_ptr()))
    return loss, gs, gt


def reciprocal_nn(A, B, dist='dot', want_A=True, want_B=True, packed=False):
    """nn_A, nn_B (int64 CUDA tensors or None) for fp32 CUDA descriptors A (nA, D), B (nB, D).

    packed=True (both directions wanted): one (nA + nB,) tensor [nn_A | nn_B], so that a caller that needs the indices
    on the host issues a single copy."""
    if dist not in DIST:
        raise ValueError(f'Unknown {dist=}')
    require_cuda(A, B)
    lib = load()
    A = A.contiguous().float()
    B = B.contiguous().float()
    nA, nB = A.shape[0], B.shape[0]
    if A.shape[1] != B.shape[1]:
        raise ValueError('descriptor dimensions differ')
    both = None
    if packed and want_A and want_B:
        both = torch.empty(nA + nB, dtype=torch.int64, device=A.device)
        nn_A, nn_B = both[:nA], both[nA:]
    else:
        nn_A = torch.empty(nA, dtype=torch.int64, device=A.device) if want_A else None
        nn_B = torch.empty(nB, dtype=torch.int64, device=A.device) if want_B else None
    ws_bytes = lib.gd3_reciprocal_nn_workspace(nA, nB)
    ws = workspace(ws_bytes, A.device)
    with torch.cuda.device(A.device):
        check(lib.gd3_reciprocal_nn(ptr(A), nA, ptr(B), nB, A.shape[1], DIST[dist], ptr(nn_A), ptr(nn_B),
                                    ptr(ws), ws.numel(), stream_ptr()))
    return both if both is not None else (nn_A, nn_B)


def fast_reciprocal_nn(pts1, pts2, seeds, max_iter=10, dist='dot', host_poll=True):
    """Device-resident reciprocal-NN ping-pong.  pts1 (n1, D), pts2 (n2, D) fp32 CUDA, seeds (S,) int32 CUDA flat
    indices into pts1 -> (xy1, xy2) int32 CUDA and converged (S,) bool CUDA.  host_poll=False never synchronises;
    host_poll=True reads the live-seed count back once per round (from the second on) to stop early."""
    if dist not in DIST:
        raise ValueError(f'Unknown {dist=}')
    require_cuda(pts1, pts2, seeds)
    lib = load()
    pts1 = pts1.contiguous().float()
    pts2 = pts2.contiguous().float()
    seeds = seeds.contiguous().to(torch.int32)
    if pts1.shape[1] != pts2.shape[1]:
        raise ValueError('descriptor dimensions differ')
    S = seeds.numel()
    xy1 = torch.empty(S, dtype=torch.int32, device=pts1.device)
    xy2 = torch.empty(S, dtype=torch.int32, device=pts1.device)
    conv = torch.empty(S, dtype=torch.uint8, device=pts1.device)
    ws = workspace(lib.gd3_fast_reciprocal_nn_workspace(S, pts1.shape[1]), pts1.device)
    with torch.cuda.device(pts1.device):
        check(lib.gd3_fast_reciprocal_nn(ptr(pts1), pts1.shape[0], ptr(pts2), pts2.shape[0], pts1.shape[1], DIST[dist],
                                         ptr(seeds), S, int(max_iter), int(bool(host_poll)), ptr(xy1), ptr(xy2), ptr(conv),
                                         ptr(ws),
                                         ws.numel(), stream_ptr()))
    return xy1, xy2, conv.bool()


def kp_prepare(kp, H, W, patch_size=None, depth=None, window=3):
    """Patch masks and / or window depths for (P, K, 2) CUDA keypoints in one launch.

    patch_size given -> bool (P, (H // p) * (W // p)) masks; depth given ((H, W) shared or (P, H, W)) -> (P, K) depths."""
    require_cuda(kp)
    lib = load()
    kp = kp.to(torch.float32).contiguous()
    P, K = kp.shape[0], kp.shape[1]
    mask = None
    if patch_size is not None:
        mask = torch.empty(P, (H // patch_size) * (W // patch_size), dtype=torch.uint8, device=kp.device)
    kd = None
    stride = 0
    if depth is not None:
        depth = depth.to(torch.float32).contiguous()
        assert depth.shape[-2:] == (H, W)
        stride = 0 if depth.dim() == 2 else H * W
        kd = torch.empty(P, K, dtype=torch.float32, device=kp.device)
    with torch.cuda.device(kp.device):
        check(lib.gd3_kp_prepare(ptr(kp), P, K, int(H), int(W), int(patch_size or 1), int(window), ptr(depth), stride,
                                 ptr(mask), ptr(kd), stream_ptr()))
    return (mask.bool() if mask is not None else None), kd


def point_cloud_to_depth(points, intrinsics, w, h):
    """(B, M, 3) CUDA camera-frame points + (3, 3) or (B, 3, 3) intrinsics -> (B, h, w) fp32 mean-z depth images."""
    require_cuda(points, intrinsics)
    lib = load()
    points = points.to(torch.float32).contiguous()
    intrinsics = intrinsics.to(torch.float32).contiguous()
    B, M = points.shape[0], points.shape[1]
    assert points.shape[2] == 3 and intrinsics.shape[-2:] == (3, 3)
    assert intrinsics.dim() == 2 or intrinsics.shape[0] == B
    depth = torch.empty(B, int(h), int(w), dtype=torch.float32, device=points.device)
    nbytes = lib.gd3_point_cloud_to_depth_workspace(B, int(h), int(w))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        check(lib.gd3_point_cloud_to_depth(ptr(points), B, M, ptr(intrinsics), 0 if intrinsics.dim() == 2 else 9,
                                           int(w), int(h), ptr(depth), ptr(ws), ws.numel(), stream_ptr()))
    return depth


def vggt_attn_accumulate(q_scaled, k, temperature=1.0, weight=1.0, out=None, skip=5, round_bf16=True):
    """One VGGT global block: (B, heads, n_tokens, head_dim) bf16 CUDA q * scale and k -> head-mean cross-view attention
    maps (attn12, attn21), each (B, n, n) fp32 with n = n_tokens // 2 - skip.  ``out=(attn12, attn21)`` adds
    ``weight`` x this block to existing maps."""
    require_cuda(q_scaled, k)
    lib = load()
    if q_scaled.dtype != torch.bfloat16 or k.dtype != torch.bfloat16:
        raise ValueError('vggt_attn_accumulate: q and k must be bfloat16 (the teacher runs under bf16 autocast)')
    if q_scaled.dim() != 4 or q_scaled.shape != k.shape:
        raise ValueError(f'vggt_attn_accumulate: q / k must both be (B, heads, tokens, head_dim), got '
                         f'{tuple(q_scaled.shape)} {tuple(k.shape)}')
    q_scaled, k = q_scaled.contiguous(), k.contiguous()
    B, heads, T, dh = q_scaled.shape
    n = T // 2 - int(skip)
    if n <= 0:
        raise ValueError(f'vggt_attn_accumulate: no patch tokens left ({T} tokens, skip {skip})')
    if out is None:
        a12 = torch.empty(B, n, n, dtype=torch.float32, device=q_scaled.device)
        a21 = torch.empty(B, n, n, dtype=torch.float32, device=q_scaled.device)
        accumulate = 0
    else:
        a12, a21 = out
        assert a12.shape == a21.shape == (B, n, n) and a12.is_contiguous() and a21.is_contiguous()
        assert a12.dtype == a21.dtype == torch.float32
        accumulate = 1
    ws = workspace(lib.gd3_vggt_attn_workspace(B, heads, max(n, 0)), q_scaled.device)
    with torch.cuda.device(q_scaled.device):
        check(lib.gd3_vggt_attn_accumulate(ptr(q_scaled), ptr(k), B, heads, T, dh, int(skip), float(temperature),
                                           int(bool(round_bf16)), float(weight), accumulate, ptr(a12), ptr(a21), ptr(ws),
                                           ws.numel(), stream_ptr()))
    return a12, a21


def teacher_volume(tgt_camap, src_camap=None, temperature=3.0, reciprocity=True, plain_mean=False):
    """tgt_attn_map (B, N, N) from the per-layer cross-attention logits (lists of (B, H, N, N) fp32 CUDA tensors).
    plain_mean=True: mean over layers and heads only (the VGGT teacher's aggregation)."""
    reciprocity = bool(reciprocity) and not plain_mean
    import ctypes
    lib = load()
    tgt = [t.contiguous().float() for t in tgt_camap]
    require_cuda(*tgt)
    src = [t.contiguous().float() for t in src_camap] if reciprocity else []
    L = len(tgt)
    B, H, N, N2 = tgt[0].shape
    assert N == N2 and all(t.shape == tgt[0].shape for t in tgt + src) and (not reciprocity or len(src) == L)
    arr_t = (ctypes.c_void_p * L)(*[t.data_ptr() for t in tgt])
    arr_s = (ctypes.c_void_p * L)(*[t.data_ptr() for t in src]) if reciprocity else None
    out = torch.empty(B, N, N, dtype=torch.float32, device=tgt[0].device)
    ws = workspace(lib.gd3_teacher_volume_workspace(L, B, N), out.device)
    with torch.cuda.device(out.device):
        check(lib.gd3_teacher_volume(arr_t, arr_s, L, B, H, N, float(temperature), 2 if plain_mean else int(reciprocity), ptr(out),
                                     ptr(ws), ws.numel(), stream_ptr()))
    return out


def semantic_argmax(kp_desc, desc2, img_size, patch_size=14, stride=14, want_val=False):
    """Arg-max pixel (y * img_size + x) of image 2 for every keypoint descriptor of image 1.

    kp_desc: (K, C) or the reference's (1, C, K) fp32 CUDA (any strides); desc2: (1, C, ph, pw) or (C, ph, pw) fp32 CUDA.
    Returns int64 (K,) [and the similarity attained]."""
    require_cuda(kp_desc, desc2)
    lib = load()
    if kp_desc.dim() == 3:                      # (1, C, K) as produced by interpolate_features
        assert kp_desc.shape[0] == 1
        kd = kp_desc[0].float()
        sk, sc = kd.stride(1), kd.stride(0)
        K, C = kd.shape[1], kd.shape[0]
    else:
        kd = kp_desc.float()
        sk, sc = kd.stride(0), kd.stride(1)
        K, C = kd.shape
    d2 = desc2.float()
    if d2.dim() == 4:
        assert d2.shape[0] == 1
        d2 = d2[0]
    d2 = d2.contiguous()
    assert d2.shape[0] == C
    ph, pw = d2.shape[1], d2.shape[2]
    idx = torch.empty(K, dtype=torch.int64, device=d2.device)
    val = torch.empty(K, dtype=torch.float32, device=d2.device) if want_val else None
    ws = workspace(lib.gd3_semantic_argmax_workspace(K, ph, pw), d2.device)
    with torch.cuda.device(d2.device):
        check(lib.gd3_semantic_argmax(ptr(kd), sk, sc, ptr(d2), K, C, ph, pw, int(img_size), int(patch_size), int(stride),
                                      ptr(idx), ptr(val), ptr(ws), ws.numel(), stream_ptr()))
    return (idx, val) if want_val else idx


def debug_gemm_bf16(A, B, tile_n=256):
    """C[b] = A[b] @ B[b].T through the tcgen05 GEMM (A: (b, M, K) bf16, B: (b, N, K) bf16) -> fp32."""
    require_cuda(A, B)
    lib = load()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.dim() == 3 and B.dim() == 3
    A = A.contiguous()
    B = B.contiguous()
    b, M, K = A.shape
    N = B.shape[1]
    C = torch.empty(b, M, N, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib.gd3_debug_gemm_bf16(ptr(A), ptr(B), ptr(C), M, N, K, b, K, K, N, tile_n, stream_ptr()))
    return C


def debug_gemm_bf16_mn(A, B, a_mn=False, b_mn=False, tile_n=256):
    """C[b] = op(A[b]) @ op(B[b]).T with MN-major operands: a_mn -> A is (b, K, M), b_mn -> B is (b, K, N)."""
    require_cuda(A, B)
    lib = load()
    A = A.contiguous()
    B = B.contiguous()
    b = A.shape[0]
    K, M = (A.shape[1], A.shape[2]) if a_mn else (A.shape[2], A.shape[1])
    N = B.shape[2] if b_mn else B.shape[1]
    C = torch.empty(b, M, N, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib.gd3_debug_gemm_bf16_mn(ptr(A), ptr(B), ptr(C), M, N, K, b, int(a_mn), int(b_mn), tile_n, stream_ptr()))
    return C


def launch_count():
    """Number of CUDA kernels lib3dgd has launched in this process."""
    return int(load().gd3_launch_count())


def profile_enable(on=True):
    load().gd3_profile_enable(int(bool(on)))


def profile_read():
    """{kernel name: (launches, total_ms)} since the last read (synchronises the device)."""
    lib = load()
    n = lib.gd3_profile_read(None, 0)
    buf = ctypes.create_string_buffer(int(n) + 64)
    # records were consumed by the sizing call only if a buffer was given; call again with the buffer
    lib.gd3_profile_read(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(' ', 2)
        out[name] = (int(cnt), float(ms))
    return out
