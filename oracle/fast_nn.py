"""Oracle restatement of ``mast3r/fast_nn.py`` (reciprocal nearest-neighbour matcher).

Test infrastructure only (see ``oracle/__init__.py``).  Runs on CPU tensors.
The scipy-KDTree branch of the reference (CPU device without ``dist`` /
``block_size``) is restated too, because the oracle is the CPU baseline.
"""

import math

import numpy as np
import torch


def _as_tensor(x):
    return torch.from_numpy(x) if isinstance(x, np.ndarray) else x


@torch.no_grad()
def bruteforce_reciprocal_nns(A, B, device='cpu', block_size=None, dist='l2'):
    """Row- and column-wise nearest neighbours of the |A| x |B| distance matrix.

    Follows ``mast3r/fast_nn.py:16-70``.  ``dist='dot'`` maximises A @ B.T,
    ``dist='l2'`` minimises ``torch.cdist``.  Ties resolve to the lowest index
    (``torch.max`` / ``torch.min`` semantics); the blocked path, taken when
    |A|*|B| > block_size**2, keeps the earlier block on ties (strict ``<``,
    ``:60-61``).  Returns two int64 numpy arrays.
    """
    A = _as_tensor(A).to(device)
    B = _as_tensor(B).to(device)
    if dist == 'l2':
        def score(x, y):
            return torch.cdist(x, y)
    elif dist == 'dot':
        def score(x, y):
            return -(x @ y.T)
    else:
        raise ValueError(f'Unknown {dist=}')

    nA, nB = A.shape[0], B.shape[0]
    if block_size is None or nA * nB <= block_size ** 2:
        d = score(A, B)
        nn_A = d.min(dim=1).indices
        nn_B = d.min(dim=0).indices
    else:
        best_A = torch.full((nA,), float('inf'), dtype=A.dtype, device=device)
        best_B = torch.full((nB,), float('inf'), dtype=B.dtype, device=device)
        nn_A = torch.full((nA,), -1, dtype=torch.int64, device=device)
        nn_B = torch.full((nB,), -1, dtype=torch.int64, device=device)
        for ia in range(math.ceil(nA / block_size)):
            ra = slice(ia * block_size, (ia + 1) * block_size)
            for ib in range(math.ceil(nB / block_size)):
                rb = slice(ib * block_size, (ib + 1) * block_size)
                d = score(A[ra], B[rb])
                va, ja = d.min(dim=1)
                vb, jb = d.min(dim=0)
                upd_a = va < best_A[ra]
                upd_b = vb < best_B[rb]
                best_A[ra] = torch.where(upd_a, va, best_A[ra])
                best_B[rb] = torch.where(upd_b, vb, best_B[rb])
                nn_A[ra] = torch.where(upd_a, ja + ib * block_size, nn_A[ra])
                nn_B[rb] = torch.where(upd_b, jb + ia * block_size, nn_B[rb])
    return nn_A.cpu().numpy(), nn_B.cpu().numpy()


class cdistMatcher:
    """Follows ``mast3r/fast_nn.py:73-84``."""

    def __init__(self, db_pts, device='cpu'):
        self.db_pts = db_pts.to(device)
        self.device = device

    def query(self, queries, k=1, **kw):
        assert k == 1
        if queries.numel() == 0:
            return None, []
        nnA, _ = bruteforce_reciprocal_nns(queries, self.db_pts, device=self.device, **kw)
        return None, nnA


def merge_corres(idx1, idx2, shape1=None, shape2=None, ret_xy=True, ret_index=False):
    """Unique (idx1, idx2) pairs sorted by idx1 then idx2, optionally as (x, y).

    Follows ``mast3r/fast_nn.py:87-106`` (which packs the int32 pair into one
    int64 with idx1 in the high word on little-endian hosts and calls np.unique).
    """
    assert idx1.dtype == idx2.dtype == np.int32
    key = (idx1.astype(np.int64) << 32) | (idx2.astype(np.int64) & 0xFFFFFFFF)
    if ret_index:
        key, first = np.unique(key, return_index=True)
    else:
        key = np.unique(key)
    u1 = (key >> 32).astype(np.int32)
    u2 = (key & 0xFFFFFFFF).astype(np.int32)
    if ret_xy:
        assert shape1 and shape2
        y1, x1 = np.unravel_index(u1, shape1)
        y2, x2 = np.unravel_index(u2, shape2)
        if ret_xy == 'y_x':
            u1, u2 = (y1, x1), (y2, x2)
        else:
            u1 = np.stack([x1, y1], axis=-1)
            u2 = np.stack([x2, y2], axis=-1)
    if ret_index:
        return u1, u2, first
    return u1, u2


def fast_reciprocal_NNs(pts1, pts2, subsample_or_initxy1=8, ret_xy=True, pixel_tol=0, ret_basin=False,
                        device='cpu', **matcher_kw):
    """Iterative reciprocal NN from a sparse seed grid.  Follows ``mast3r/fast_nn.py:109-188``."""
    H1, W1, D1 = pts1.shape
    H2, W2, D2 = pts2.shape
    assert D1 == D2
    pts1 = pts1.reshape(-1, D1)
    pts2 = pts2.reshape(-1, D2)

    if isinstance(subsample_or_initxy1, int) and pixel_tol == 0:
        S = subsample_or_initxy1
        y1, x1 = np.mgrid[S // 2:H1:S, S // 2:W1:S].reshape(2, -1)
        max_iter = 10
    else:
        x1, y1 = subsample_or_initxy1
        x1 = x1.cpu().numpy() if isinstance(x1, torch.Tensor) else x1
        y1 = y1.cpu().numpy() if isinstance(y1, torch.Tensor) else y1
        max_iter = 1

    xy1 = np.int32(np.unique(x1 + W1 * y1))
    xy2 = np.full_like(xy1, -1)
    old_xy1, old_xy2 = xy1.copy(), xy2.copy()

    is_cuda = (isinstance(device, str) and device.startswith('cuda')) or \
              (isinstance(device, torch.device) and device.type.startswith('cuda'))
    if 'dist' in matcher_kw or 'block_size' in matcher_kw or is_cuda:
        pts1, pts2 = pts1.to(device), pts2.to(device)
        tree1, tree2 = cdistMatcher(pts1, device=device), cdistMatcher(pts2, device=device)
    else:
        from scipy.spatial import KDTree
        pts1 = pts1.cpu().numpy() if isinstance(pts1, torch.Tensor) else pts1
        pts2 = pts2.cpu().numpy() if isinstance(pts2, torch.Tensor) else pts2
        tree1, tree2 = KDTree(pts1), KDTree(pts2)

    notyet = np.ones(len(xy1), dtype=bool)
    basin = np.full((H1 * W1 + 1,), -1, dtype=np.int32) if ret_basin else None
    it = 0
    while notyet.any():
        _, nn = tree2.query(pts1[xy1[notyet]], **matcher_kw)
        xy2[notyet] = np.asarray(nn)
        if not ret_basin:
            notyet &= (old_xy2 != xy2)
        _, nn = tree1.query(pts2[xy2[notyet]], **matcher_kw)
        xy1[notyet] = np.asarray(nn)
        if ret_basin:
            basin[old_xy1[notyet]] = xy1[notyet]
        notyet &= (old_xy1 != xy1)
        it += 1
        if it >= max_iter:
            break
        old_xy2[:] = xy2
        old_xy1[:] = xy1

    if pixel_tol > 0:
        old_yx = np.stack(np.unravel_index(old_xy1, (H1, W1)), axis=-1)
        new_yx = np.stack(np.unravel_index(xy1, (H1, W1)), axis=-1)
        converged = np.linalg.norm(old_yx - new_yx, axis=-1) < pixel_tol
        if not isinstance(subsample_or_initxy1, int):
            xy1 = old_xy1
    else:
        converged = ~notyet

    out1, out2 = merge_corres(xy1[converged], xy2[converged], (H1, W1), (H2, W2), ret_xy=ret_xy)
    if ret_basin:
        return out1, out2, basin
    return out1, out2


def extract_correspondences_nonsym(A, B, confA, confB, subsample=8, device='cpu', ptmap_key='pred_desc', pixel_tol=0):
    """Matches searched from both images, merged, with min(confA, confB) per match.

    Follows ``mast3r/fast_nn.py:191-223``.  Returns (xy1, xy2, conf) as torch tensors like the
    reference's ``todevice``.
    """
    if '3d' in ptmap_key:
        opt = dict(device='cpu', workers=32)
    else:
        opt = dict(device=device, dist='dot', block_size=2 ** 13)
    HA, WA = A.shape[:2]
    HB, WB = B.shape[:2]
    if pixel_tol == 0:
        ab = fast_reciprocal_NNs(A, B, subsample_or_initxy1=subsample, ret_xy=False, **opt)
        ba = fast_reciprocal_NNs(B, A, subsample_or_initxy1=subsample, ret_xy=False, **opt)
    else:
        yA, xA = np.mgrid[subsample // 2:HA:subsample, subsample // 2:WA:subsample].reshape(2, -1)
        yB, xB = np.mgrid[subsample // 2:HB:subsample, subsample // 2:WB:subsample].reshape(2, -1)
        ab = fast_reciprocal_NNs(A, B, subsample_or_initxy1=(xA, yA), ret_xy=False, pixel_tol=pixel_tol, **opt)
        ba = fast_reciprocal_NNs(B, A, subsample_or_initxy1=(xB, yB), ret_xy=False, pixel_tol=pixel_tol, **opt)
    idx1 = np.r_[ab[0], ba[1]]
    idx2 = np.r_[ab[1], ba[0]]
    c1 = np.asarray(confA).ravel()[idx1]
    c2 = np.asarray(confB).ravel()[idx2]
    xy1, xy2, first = merge_corres(idx1, idx2, (HA, WA), (HB, WB), ret_xy=True, ret_index=True)
    conf = np.minimum(c1[first], c2[first])
    return tuple(torch.from_numpy(np.ascontiguousarray(v)) for v in (xy1.copy(), xy2.copy(), conf))
