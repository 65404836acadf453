"""numpy oracle of the device-side resampler (TEST INFRASTRUCTURE ONLY) and the reference's loader semantics.

``resample_indices`` restates csrc/data.cu: the replacement for sample_pts (dataloader_multi_peak_vonMises.py:21-26:
``arr[np.random.choice(n, num, replace=(n < num))]``) - a uniform subset without replacement when n >= num (the `num`
smallest pseudo-random keys, ties to the lower index, ascending order), independent uniform draws with replacement
otherwise.  The index streams are this implementation's own (numpy's legacy generator is not replayed), so the parity
with the reference is distributional: tests check the subset / marginal statistics against np.random.choice.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def resample_indices(n: int, num: int, seed: int, draw: int) -> np.ndarray:
    """Source rows (num,) int64 chosen for a cloud of n points by draw number `draw` (= base_draw + counter * B + b)."""
    with np.errstate(over="ignore"):
        k0 = mix64(np.uint64(seed) ^ mix64(np.uint64(draw)))
        if n <= 0:
            return np.full(num, -1, dtype=np.int64)
        if n < num:
            r = mix64(k0 + np.arange(num, dtype=np.uint64))
            return np.array([(int(v) * n) >> 64 for v in r], dtype=np.int64)        # floor(u * n), u = r / 2^64
        key = (mix64(k0 + np.arange(n, dtype=np.uint64)) >> np.uint64(32)).astype(np.uint64)
    order = np.lexsort((np.arange(n), key))                                       # by key, ties by index
    return np.sort(order[:num]).astype(np.int64)
