"""numpy oracle of centroid sampling and neighbour grouping (integer outputs, bit-exact contracts)."""
from __future__ import annotations

import numpy as np


def farthest_point_sample(xyz: np.ndarray, npoint: int, start_idx: np.ndarray) -> np.ndarray:
    """FPS indices (B,npoint) int64.  Follows farthest_point_sample, PointNet++Demo.py:8-29:
    running minimum initialised to 1e10 (:19), first centroid = start_idx (the reference draws it
    with torch.randint, :20), dist = sum((xyz - centroid)**2, -1) in fp32 (:25) which evaluates as
    ((dx*dx)+(dy*dy))+(dz*dz) with every operation rounded, update where dist < distance (:26-27),
    next = argmax with first-occurrence ties (:28)."""
    xyz = np.asarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.zeros((B, npoint), dtype=np.int64)
    for b in range(B):
        p = xyz[b]
        dmin = np.full(N, 1e10, dtype=np.float32)
        far = int(start_idx[b])
        for i in range(npoint):
            out[b, i] = far
            d = p - p[far]                      # fp32 subtraction
            sq = d * d                          # fp32 squares
            dist = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
            m = dist < dmin
            dmin[m] = dist[m]
            far = int(np.argmax(dmin))          # first occurrence of the maximum
    return out


def ball_query(radius: float, nsample: int, xyz: np.ndarray, new_xyz: np.ndarray) -> np.ndarray:
    """(B,S,nsample) int64.  Follows query_ball_point, PointNet++Demo.py:49-70: fp32 direct squared
    distances (:63), points with d2 > float32(radius**2) excluded (:65), ascending index order, first
    nsample (:66), short rows padded with the row's first hit (:67-69); an empty row is all N."""
    xyz = np.asarray(xyz, dtype=np.float32)
    new_xyz = np.asarray(new_xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    r2 = np.float32(float(radius) ** 2)
    out = np.empty((B, S, nsample), dtype=np.int64)
    for b in range(B):
        d = new_xyz[b][:, None, :] - xyz[b][None, :, :]
        sq = d * d
        dist = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        for s in range(S):
            hits = np.nonzero(~(dist[s] > r2))[0]
            row = np.full(nsample, N, dtype=np.int64)
            n = min(nsample, hits.size)
            row[:n] = hits[:n]
            if hits.size:
                row[n:] = hits[0]
            out[b, s] = row
    return out


def square_distance(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """(B,N,M) fp32.  Follows square_distance, models/base.py:20-27: -2 src @ dst^T, + sum(src^2), + sum(dst^2), in
    that order, every step in fp32 (the matmul's internal summation order is the BLAS's: compare to ~1e-6)."""
    src = np.asarray(src, dtype=np.float32)
    dst = np.asarray(dst, dtype=np.float32)
    d = np.float32(-2.0) * np.matmul(src, dst.transpose(0, 2, 1))
    d = d + (src ** 2).sum(-1)[:, :, None]
    d = d + (dst ** 2).sum(-1)[:, None, :]
    return d.astype(np.float32)


def knn(new_xyz: np.ndarray, xyz: np.ndarray, nsample: int):
    """k nearest neighbours as SETS.  Follows query_ball_point, models/base.py:29-35
    (square_distance :20-27 + topk(largest=False, sorted=False)).  The reference's row order is
    unspecified and its fp32 `-2ab+a^2+b^2` bits depend on the BLAS, so the oracle is the exact
    answer: fp64 direct distances, stable argsort.  Returns (idx sorted ascending per row,
    rel_margin) where rel_margin[b,s] = (d[k] - d[k-1]) / d[k] is the gap between the last kept and
    the first rejected neighbour: rows with a margin below fp32 resolution are ties on which any
    fp32 implementation (including the reference on another BLAS) may legitimately differ."""
    xyz64 = np.asarray(xyz, dtype=np.float64)
    new64 = np.asarray(new_xyz, dtype=np.float64)
    B, N, _ = xyz64.shape
    S = new64.shape[1]
    idx = np.empty((B, S, nsample), dtype=np.int64)
    margin = np.full((B, S), np.inf)
    for b in range(B):
        d = ((new64[b][:, None, :] - xyz64[b][None, :, :]) ** 2).sum(-1)
        order = np.argsort(d, axis=1, kind="stable")
        idx[b] = np.sort(order[:, :nsample], axis=1)
        if nsample < N:
            dk = np.take_along_axis(d, order[:, nsample - 1:nsample + 1], axis=1)
            margin[b] = (dk[:, 1] - dk[:, 0]) / np.maximum(dk[:, 1], 1e-30)
    return idx, margin


def knn_rows_match(got: np.ndarray, want_sorted: np.ndarray, margin: np.ndarray, tol: float = 2e-6):
    """Set equality per row, except rows whose k-th/k+1-th gap is below `tol` (fp32 near-ties).
    Returns (n_rows, n_equal, n_tie_rows_excused, n_bad)."""
    got_sorted = np.sort(np.asarray(got, dtype=np.int64), axis=-1)
    eq = (got_sorted == want_sorted).all(-1)
    tie = margin < tol
    bad = ~eq & ~tie
    return eq.size, int(eq.sum()), int((~eq & tie).sum()), int(bad.sum())


def randperm_subset_replay(seed: int, B: int, N: int, npoint: int):
    """The reference's "fps_idx": torch.stack([torch.randperm(N)[:npoint] for _ in range(B)]) on the
    CPU generator (models/pointnet_pp_8dir.py:28) after torch.manual_seed(seed)."""
    import torch
    torch.manual_seed(seed)
    return torch.stack([torch.randperm(N)[:npoint] for _ in range(B)]).numpy()
