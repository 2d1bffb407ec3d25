"""CUDA drop-in for the reference's ``mast3r/fast_nn.py``: same public names, arguments, return types and errors.

What runs where
* every nearest-neighbour search is ``gd3_reciprocal_nn`` (fused scores + row / column arg-best, lowest index on ties
  like ``torch.max`` / ``torch.min``; the score matrix is never materialised);
* the seed ping-pong of ``fast_reciprocal_NNs`` (``mast3r/fast_nn.py:147-170``) is ``gd3_fast_reciprocal_nn``: the seed
  state stays on the GPU for all rounds and comes back once;
* only ``ret_basin=True`` keeps a host loop, because the basin table is host bookkeeping in the reference as well
  (``:172-176``); unique + sort (``merge_corres``) is numpy as in the reference.

Not provided: the scipy-KDTree branch the reference takes for CPU devices without ``dist`` / ``block_size``
(``mast3r/fast_nn.py:142-145``) -- there is no CPU path in this library.
"""
import numpy as np
import torch

from .. import _lib

_DISTS = ('l2', 'dot')


def _is_cuda_device(device):
    if isinstance(device, torch.device):
        return device.type == 'cuda'
    return isinstance(device, str) and device.startswith('cuda')


def _need_gpu(device):
    if not _is_cuda_device(device):
        raise _lib.Gd3Error(f'gd3 fast_nn has no CPU path (device={device!r})')
    _lib.require_cuda()


def _on_device(x, device):
    return (torch.from_numpy(x) if isinstance(x, np.ndarray) else x).to(device)


def _matcher_options(kw):
    """Validate the reference's ``**matcher_kw`` (dist / block_size) and return the distance name."""
    extra = set(kw) - {'dist', 'block_size'}
    if extra:
        raise TypeError(f'unexpected matcher arguments {sorted(extra)}')
    dist = kw.get('dist', 'l2')
    if dist not in _DISTS:
        raise ValueError(f'Unknown {dist=}')
    return dist


@torch.no_grad()
def bruteforce_reciprocal_nns(A, B, device='cuda', block_size=None, dist='l2'):
    """``mast3r/fast_nn.py:16-70``: (nn_A, nn_B) int64 numpy arrays.

    ``block_size`` only bounded the reference's temporary and never changed the result; the fused kernel has no
    temporary, so the argument is accepted and ignored.
    """
    if dist not in _DISTS:
        raise ValueError(f'Unknown {dist=}')
    _need_gpu(device)
    A, B = _on_device(A, device), _on_device(B, device)
    # one device -> host copy (and one synchronisation) for both index arrays ("one D2H + sync at the end")
    both = _lib.reciprocal_nn(A, B, dist=dist, packed=True).cpu().numpy()
    return both[:A.shape[0]], both[A.shape[0]:]


class cdistMatcher:
    """``mast3r/fast_nn.py:73-84``: a brute-force "tree" over device-resident points; ``query`` returns (None, nn)."""

    def __init__(self, db_pts, device='cuda'):
        _need_gpu(device)
        self.device = device
        self.db_pts = db_pts.to(device).contiguous().float()

    def query(self, queries, k=1, **kw):
        assert k == 1
        if queries.numel() == 0:
            return None, []
        dist = _matcher_options(kw)
        nn, _ = _lib.reciprocal_nn(_on_device(queries, self.device), self.db_pts, dist=dist, want_B=False)
        return None, nn.cpu().numpy()


def merge_corres(idx1, idx2, shape1=None, shape2=None, ret_xy=True, ret_index=False):
    """``mast3r/fast_nn.py:87-106``: unique (idx1, idx2) pairs ordered by idx1 then idx2, optionally as pixel (x, y)."""
    assert idx1.dtype == idx2.dtype == np.int32
    key = (idx1.astype(np.int64) << 32) | (idx2.astype(np.int64) & 0xFFFFFFFF)     # sorts like the pair
    if ret_index:
        key, first = np.unique(key, return_index=True)
    else:
        key = np.unique(key)
    xy1 = (key >> 32).astype(np.int32)
    xy2 = (key & 0xFFFFFFFF).astype(np.int32)
    if ret_xy:
        assert shape1 and shape2
        rc1 = np.unravel_index(xy1, shape1)
        rc2 = np.unravel_index(xy2, shape2)
        if ret_xy == 'y_x':
            xy1, xy2 = rc1, rc2
        else:
            xy1 = np.stack((rc1[1], rc1[0]), axis=-1)
            xy2 = np.stack((rc2[1], rc2[0]), axis=-1)
    return (xy1, xy2, first) if ret_index else (xy1, xy2)


def _seed_indices(spec, H1, W1, grid_mode):
    """Flat, unique, sorted int32 seed indices into image 1 (``:116-131``)."""
    if grid_mode:
        ys, xs = np.mgrid[spec // 2:H1:spec, spec // 2:W1:spec].reshape(2, -1)
    else:
        xs, ys = spec
        xs = xs.cpu().numpy() if isinstance(xs, torch.Tensor) else xs
        ys = ys.cpu().numpy() if isinstance(ys, torch.Tensor) else ys
    return np.int32(np.unique(xs + W1 * ys))


def _basin_walk(pts1, pts2, seeds, H1, W1, rounds, dist):
    """The ``ret_basin=True`` variant of the ping-pong (``:147-176``): seeds are not retired on the 1 -> 2 leg and every
    round records where each live seed of image 1 moved to.  Host loop (the basin table is host data)."""
    cur1, cur2 = seeds.copy(), np.full_like(seeds, -1)
    prev1 = cur1.copy()
    live = np.ones(len(seeds), dtype=bool)
    basin = np.full((H1 * W1 + 1,), -1, dtype=np.int32)

    def nearest(queries, db):
        nn, _ = _lib.reciprocal_nn(queries, db, dist=dist, want_B=False)
        return nn.cpu().numpy()

    def rows(pts, idx):
        return pts[torch.from_numpy(idx.astype(np.int64)).to(pts.device)]

    for r in range(rounds):
        if not live.any():
            break
        cur2[live] = nearest(rows(pts1, cur1[live]), pts2)
        cur1[live] = nearest(rows(pts2, cur2[live]), pts1)
        basin[prev1[live]] = cur1[live]
        live &= prev1 != cur1
        if r + 1 < rounds:
            prev1[:] = cur1
    return cur1, cur2, ~live, basin


def fast_reciprocal_NNs(pts1, pts2, subsample_or_initxy1=8, ret_xy=True, pixel_tol=0, ret_basin=False,
                        device='cuda', **matcher_kw):
    """``mast3r/fast_nn.py:109-188``: iterative reciprocal nearest neighbours from a sparse seed grid (10 rounds) or from
    explicit seeds (one round)."""
    H1, W1, DIM1 = pts1.shape
    H2, W2, DIM2 = pts2.shape
    assert DIM1 == DIM2
    if not ('dist' in matcher_kw or 'block_size' in matcher_kw or _is_cuda_device(device)):
        raise _lib.Gd3Error('the scipy-KDTree CPU branch of fast_reciprocal_NNs is not provided '
                            '(pass device="cuda" or dist=/block_size=)')
    _need_gpu(device)
    dist = _matcher_options(matcher_kw)
    flat1 = _on_device(pts1, device).reshape(-1, DIM1).contiguous().float()
    flat2 = _on_device(pts2, device).reshape(-1, DIM2).contiguous().float()

    grid_mode = isinstance(subsample_or_initxy1, int) and pixel_tol == 0
    rounds = 10 if grid_mode else 1
    seeds = _seed_indices(subsample_or_initxy1, H1, W1, grid_mode)

    if ret_basin:
        xy1, xy2, converged, basin = _basin_walk(flat1, flat2, seeds, H1, W1, rounds, dist)
    else:
        basin = None
        if len(seeds) == 0:
            xy1 = xy2 = np.zeros(0, dtype=np.int32)
            converged = np.zeros(0, dtype=bool)
        else:
            d1, d2, conv = _lib.fast_reciprocal_nn(flat1, flat2, torch.from_numpy(seeds).to(flat1.device),
                                                   max_iter=rounds, dist=dist)
            packed = torch.stack([d1, d2, conv.to(torch.int32)]).cpu().numpy()      # the only copy back
            xy1, xy2, converged = packed[0], packed[1], packed[2].astype(bool)

    if pixel_tol > 0:
        # explicit seeds, one round: a seed counts as converged when it came back within pixel_tol of where it
        # started, and the reported position in image 1 is the seed itself (:172-180)
        start = np.stack(np.unravel_index(seeds, (H1, W1)), axis=-1)
        back = np.stack(np.unravel_index(xy1, (H1, W1)), axis=-1)
        converged = np.linalg.norm(start - back, axis=-1) < pixel_tol
        if not isinstance(subsample_or_initxy1, int):
            xy1 = seeds

    out = merge_corres(xy1[converged], xy2[converged], (H1, W1), (H2, W2), ret_xy=ret_xy)
    return out + (basin,) if ret_basin else out


def extract_correspondences_nonsym(A, B, confA, confB, subsample=8, device=None, ptmap_key='pred_desc',
                                   pixel_tol=0):
    """``mast3r/fast_nn.py:191-223`` for descriptor maps: matches found from both sides are merged and carry the smaller
    of the two confidences.  (``'3d' in ptmap_key`` is the reference's CPU KDTree configuration: not provided.)"""
    if '3d' in ptmap_key:
        raise _lib.Gd3Error('extract_correspondences_nonsym on 3-D point maps uses the CPU KDTree branch, '
                            'which gd3 does not provide')
    search = dict(device=device, dist='dot', block_size=2 ** 13, ret_xy=False)
    shapeA, shapeB = A.shape[:2], B.shape[:2]

    def seeds_of(shape):
        if pixel_tol == 0:
            return subsample
        ys, xs = np.mgrid[subsample // 2:shape[0]:subsample, subsample // 2:shape[1]:subsample].reshape(2, -1)
        return xs, ys

    a_from_a, b_from_a = fast_reciprocal_NNs(A, B, subsample_or_initxy1=seeds_of(shapeA), pixel_tol=pixel_tol, **search)
    b_from_b, a_from_b = fast_reciprocal_NNs(B, A, subsample_or_initxy1=seeds_of(shapeB), pixel_tol=pixel_tol, **search)
    idxA = np.r_[a_from_a, a_from_b]
    idxB = np.r_[b_from_a, b_from_b]

    def flat_conf(c):
        return (c.detach().cpu().numpy() if torch.is_tensor(c) else np.asarray(c)).ravel()

    cA, cB = flat_conf(confA)[idxA], flat_conf(confB)[idxB]
    xyA, xyB, first = merge_corres(idxA, idxB, shapeA, shapeB, ret_xy=True, ret_index=True)
    conf = np.minimum(cA[first], cB[first])
    # the reference finishes with dust3r's todevice(corres, device): numpy -> torch -> device
    return tuple(torch.from_numpy(np.ascontiguousarray(v)).to(device) for v in (xyA.copy(), xyB.copy(), conf))
