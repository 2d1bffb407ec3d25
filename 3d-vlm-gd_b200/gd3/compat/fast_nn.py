"""CUDA drop-in for the reference's ``mast3r/fast_nn.py`` (same names, arguments and errors).

The score matrix is never materialised: ``gd3_reciprocal_nn`` fuses the dot / l2 scores with the
row and column arg-best (lowest index on ties, like ``torch.max`` / ``torch.min``).  Host-side
bookkeeping (seed grid, convergence masks, unique + sort) stays in numpy exactly as in the
reference, but one query costs a single small device->host copy instead of one per 8192-block.

Not provided: the scipy-KDTree branch the reference takes for CPU devices without ``dist`` /
``block_size`` (``mast3r/fast_nn.py:142-145``) -- there is no CPU path in this library.
"""
import numpy as np
import torch

from .. import _lib


def _is_cuda_device(device):
    if isinstance(device, torch.device):
        return device.type == 'cuda'
    return isinstance(device, str) and device.startswith('cuda')


def _need_gpu(device):
    if not _is_cuda_device(device):
        raise _lib.Gd3Error(f'gd3 fast_nn has no CPU path (device={device!r})')
    _lib.require_cuda()


def _to_device(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device)


@torch.no_grad()
def bruteforce_reciprocal_nns(A, B, device='cuda', block_size=None, dist='l2'):
    """Mirror of ``mast3r/fast_nn.py:16-70``.  Returns (nn_A, nn_B) as int64 numpy arrays.

    ``block_size`` is accepted for signature compatibility; it only bounded the reference's
    temporary and never changed the result (the fused kernel has no temporary).
    """
    if dist not in ('l2', 'dot'):
        raise ValueError(f'Unknown {dist=}')
    _need_gpu(device)
    A = _to_device(A, device)
    B = _to_device(B, device)
    nn_A, nn_B = _lib.reciprocal_nn(A, B, dist=dist)
    return nn_A.cpu().numpy(), nn_B.cpu().numpy()


class cdistMatcher:
    """Mirror of ``mast3r/fast_nn.py:73-84``: a brute-force 'tree' over device-resident points."""

    def __init__(self, db_pts, device='cuda'):
        _need_gpu(device)
        self.db_pts = db_pts.to(device).contiguous().float()
        self.device = device

    def query(self, queries, k=1, **kw):
        assert k == 1
        if queries.numel() == 0:
            return None, []
        dist = kw.get('dist', 'l2')
        unknown = set(kw) - {'dist', 'block_size'}
        if unknown:
            raise TypeError(f'unexpected matcher arguments {sorted(unknown)}')
        if dist not in ('l2', 'dot'):
            raise ValueError(f'Unknown {dist=}')
        nn_A, _ = _lib.reciprocal_nn(_to_device(queries, self.device), self.db_pts, dist=dist, want_B=False)
        return None, nn_A.cpu().numpy()


def merge_corres(idx1, idx2, shape1=None, shape2=None, ret_xy=True, ret_index=False):
    """Mirror of ``mast3r/fast_nn.py:87-106``: unique pairs sorted by idx1 then idx2."""
    assert idx1.dtype == idx2.dtype == np.int32
    packed = (idx1.astype(np.int64) << 32) | (idx2.astype(np.int64) & 0xFFFFFFFF)
    if ret_index:
        packed, indices = np.unique(packed, return_index=True)
    else:
        packed = np.unique(packed)
    xy1 = (packed >> 32).astype(np.int32)
    xy2 = (packed & 0xFFFFFFFF).astype(np.int32)
    if ret_xy:
        assert shape1 and shape2
        yx1 = np.unravel_index(xy1, shape1)
        yx2 = np.unravel_index(xy2, shape2)
        if ret_xy == 'y_x':
            xy1, xy2 = yx1, yx2
        else:
            xy1 = np.stack(yx1[::-1], axis=-1)
            xy2 = np.stack(yx2[::-1], axis=-1)
    if ret_index:
        return xy1, xy2, indices
    return xy1, xy2


def fast_reciprocal_NNs(pts1, pts2, subsample_or_initxy1=8, ret_xy=True, pixel_tol=0, ret_basin=False,
                        device='cuda', **matcher_kw):
    """Mirror of ``mast3r/fast_nn.py:109-188`` (iterative reciprocal NN from a sparse seed grid)."""
    H1, W1, DIM1 = pts1.shape
    H2, W2, DIM2 = pts2.shape
    assert DIM1 == DIM2

    if not ('dist' in matcher_kw or 'block_size' in matcher_kw or _is_cuda_device(device)):
        raise _lib.Gd3Error('the scipy-KDTree CPU branch of fast_reciprocal_NNs is not provided '
                            '(pass device="cuda" or dist=/block_size=)')
    _need_gpu(device)

    pts1 = _to_device(pts1, device).reshape(-1, DIM1).contiguous().float()
    pts2 = _to_device(pts2, device).reshape(-1, DIM2).contiguous().float()

    if isinstance(subsample_or_initxy1, int) and pixel_tol == 0:
        S = subsample_or_initxy1
        y1, x1 = np.mgrid[S // 2:H1:S, S // 2:W1:S].reshape(2, -1)
        max_iter = 10
    else:
        x1, y1 = subsample_or_initxy1
        if isinstance(x1, torch.Tensor):
            x1 = x1.cpu().numpy()
        if isinstance(y1, torch.Tensor):
            y1 = y1.cpu().numpy()
        max_iter = 1

    xy1 = np.int32(np.unique(x1 + W1 * y1))
    dist = matcher_kw.get('dist', 'l2')
    unknown = set(matcher_kw) - {'dist', 'block_size'}
    if unknown:
        raise TypeError(f'unexpected matcher arguments {sorted(unknown)}')
    if dist not in ('l2', 'dot'):
        raise ValueError(f'Unknown {dist=}')
    if pixel_tol == 0 and not ret_basin:
        # the whole ping-pong runs on the device (gd3_fast_reciprocal_nn): one copy back at the end instead of
        # one per query; ret_basin / pixel_tol keep the reference's host bookkeeping below
        if len(xy1) == 0:
            empty = np.zeros(0, dtype=np.int32)
            return merge_corres(empty, empty, (H1, W1), (H2, W2), ret_xy=ret_xy)
        seeds = torch.from_numpy(xy1).to(pts1.device)
        d1, d2, conv = _lib.fast_reciprocal_nn(pts1, pts2, seeds, max_iter=max_iter, dist=dist)
        keep = conv.nonzero().squeeze(1)
        out = torch.stack([d1[keep], d2[keep]]).cpu().numpy()
        return merge_corres(out[0], out[1], (H1, W1), (H2, W2), ret_xy=ret_xy)
    xy2 = np.full_like(xy1, -1)
    old_xy1 = xy1.copy()
    old_xy2 = xy2.copy()

    tree1 = cdistMatcher(pts1, device=device)
    tree2 = cdistMatcher(pts2, device=device)

    def gather(pts, idx):
        return pts[torch.from_numpy(idx.astype(np.int64)).to(pts.device)]

    notyet = np.ones(len(xy1), dtype=bool)
    basin = np.full((H1 * W1 + 1,), -1, dtype=np.int32) if ret_basin else None

    niter = 0
    while notyet.any():
        _, nn = tree2.query(gather(pts1, xy1[notyet]), **matcher_kw)
        xy2[notyet] = nn
        if not ret_basin:
            notyet &= (old_xy2 != xy2)
        _, nn = tree1.query(gather(pts2, xy2[notyet]), **matcher_kw)
        xy1[notyet] = nn
        if ret_basin:
            basin[old_xy1[notyet]] = xy1[notyet]
        notyet &= (old_xy1 != xy1)
        niter += 1
        if niter >= max_iter:
            break
        old_xy2[:] = xy2
        old_xy1[:] = xy1

    if pixel_tol > 0:
        old_yx1 = np.stack(np.unravel_index(old_xy1, (H1, W1)), axis=-1)
        new_yx1 = np.stack(np.unravel_index(xy1, (H1, W1)), axis=-1)
        converged = np.linalg.norm(old_yx1 - new_yx1, axis=-1) < pixel_tol
        if not isinstance(subsample_or_initxy1, int):
            xy1 = old_xy1
    else:
        converged = ~notyet

    xy1, xy2 = merge_corres(xy1[converged], xy2[converged], (H1, W1), (H2, W2), ret_xy=ret_xy)
    if ret_basin:
        return xy1, xy2, basin
    return xy1, xy2


def extract_correspondences_nonsym(A, B, confA, confB, subsample=8, device=None, ptmap_key='pred_desc',
                                   pixel_tol=0):
    """Mirror of ``mast3r/fast_nn.py:191-223`` for descriptor maps (``'3d' in ptmap_key`` is the
    reference's CPU KDTree configuration and is not provided)."""
    if '3d' in ptmap_key:
        raise _lib.Gd3Error('extract_correspondences_nonsym on 3-D point maps uses the CPU KDTree branch, '
                            'which gd3 does not provide')
    opt = dict(device=device, dist='dot', block_size=2 ** 13)
    HA, WA = A.shape[:2]
    HB, WB = B.shape[:2]
    if pixel_tol == 0:
        nn1to2 = fast_reciprocal_NNs(A, B, subsample_or_initxy1=subsample, ret_xy=False, **opt)
        nn2to1 = fast_reciprocal_NNs(B, A, subsample_or_initxy1=subsample, ret_xy=False, **opt)
    else:
        S = subsample
        yA, xA = np.mgrid[S // 2:HA:S, S // 2:WA:S].reshape(2, -1)
        yB, xB = np.mgrid[S // 2:HB:S, S // 2:WB:S].reshape(2, -1)
        nn1to2 = fast_reciprocal_NNs(A, B, subsample_or_initxy1=(xA, yA), ret_xy=False, pixel_tol=pixel_tol, **opt)
        nn2to1 = fast_reciprocal_NNs(B, A, subsample_or_initxy1=(xB, yB), ret_xy=False, pixel_tol=pixel_tol, **opt)

    idx1 = np.r_[nn1to2[0], nn2to1[1]]
    idx2 = np.r_[nn1to2[1], nn2to1[0]]
    confA = confA.detach().cpu().numpy() if torch.is_tensor(confA) else np.asarray(confA)
    confB = confB.detach().cpu().numpy() if torch.is_tensor(confB) else np.asarray(confB)
    c1 = confA.ravel()[idx1]
    c2 = confB.ravel()[idx2]
    xy1, xy2, idx = merge_corres(idx1, idx2, (HA, WA), (HB, WB), ret_xy=True, ret_index=True)
    conf = np.minimum(c1[idx], c2[idx])
    out = (xy1.copy(), xy2.copy(), conf)
    # the reference ends with dust3r's todevice(corres, device): numpy -> torch, then .to(device)
    return tuple(torch.from_numpy(np.ascontiguousarray(x)).to(device) for x in out)
