"""Oracle: density / alpha / bbox keep-masks (TEST INFRASTRUCTURE).

Restates data_processor.py:11-117 (density), :184-213 (alpha), :215-231 (bbox) on
plain arrays, returning the boolean mask instead of compacting.  The cluster
discovery order (Python ``set`` iteration over lexicographically inserted voxel
tuples, stable sort by size) is reproduced because it decides ties.
These are differential-tested against the imported reference in
tests/golden/make_goldens.py (the reference itself is the pin for them).
"""
from __future__ import annotations

from collections import deque
import numpy as np


def density_slider(sensitivity: float):
    """data_processor.py:17-28."""
    voxel = max(0.1, 2.0 - (sensitivity * 1.8))
    thr = 0.1 + (sensitivity * 0.9)
    return voxel, thr


def density_mask(coords: np.ndarray, voxel_size=1.0, threshold_percentage=0.32, sensitivity=None,
                 keep_multicluster=False):
    if sensitivity is not None:
        voxel_size, threshold_percentage = density_slider(sensitivity)
    n = coords.shape[0]
    q = np.floor(coords / voxel_size).astype(np.int64)
    uniq, inverse, counts = np.unique(q, axis=0, return_inverse=True, return_counts=True)
    inverse = inverse.reshape(-1)
    min_points = int(n * (threshold_percentage / 100.0))
    dense_idx = np.where(counts >= min_points)[0]
    if len(dense_idx) == 0:
        return np.zeros(n, dtype=bool), dict(clusters=0, max_len=0, dense=0)
    dense = set(map(tuple, uniq[dense_idx]))
    seen, clusters = set(), []
    for v in dense:
        if v in seen:
            continue
        comp = {v}
        seen.add(v)
        q_ = deque([v])
        while q_:
            x, y, z = q_.popleft()
            for nb in ((x - 1, y, z), (x + 1, y, z), (x, y - 1, z), (x, y + 1, z), (x, y, z - 1), (x, y, z + 1)):
                if nb in dense and nb not in seen:
                    seen.add(nb)
                    comp.add(nb)
                    q_.append(nb)
        clusters.append(comp)
    clusters.sort(key=len, reverse=True)
    max_len = len(clusters[0])
    min_size = max_len * 0.05 if keep_multicluster else max_len
    valid, kept = set(), 0
    for c in clusters:
        if len(c) >= min_size:
            valid.update(c)
            kept += 1
            if not keep_multicluster:
                break
    in_cluster = np.array([tuple(v) in valid for v in uniq])
    return in_cluster[inverse], dict(clusters=kept, max_len=max_len, dense=len(dense_idx))


def alpha_mask(opacity: np.ndarray, min_opacity_u8):
    """data_processor.py:199-208.  None == no-op (limit<=0)."""
    limit = min_opacity_u8
    if limit <= 0:
        return np.ones(opacity.shape[0], dtype=bool)
    if limit >= 255:
        return np.zeros(opacity.shape[0], dtype=bool)
    a = np.clip(limit / 255.0, 1e-6, 1.0 - 1e-6)
    t = np.log(a / (1.0 - a))
    return opacity >= t


def bbox_mask(x, y, z, min_x, min_y, min_z, max_x, max_y, max_z):
    """data_processor.py:217-224 (closed intervals; python-float bounds are NumPy-2 weak scalars)."""
    return (x >= min_x) & (x <= max_x) & (y >= min_y) & (y <= max_y) & (z >= min_z) & (z <= max_z)
