"""Oracle: Lloyd K-Means of gpu_ops.py:27-46,57-96,178-191 (TEST INFRASTRUCTURE).

Init is injectable; ``kmeans_reference_init`` makes exactly the reference's RNG
call (gpu_ops.py:182) so a caller-side ``np.random.seed`` reproduces it.
"""
from __future__ import annotations

import ctypes
import numpy as np


def kmeans_reference_init(data: np.ndarray, k: int) -> np.ndarray:
    n = data.shape[0]
    return data.astype(np.float32)[np.random.choice(n, k, replace=False)].astype(np.float32)


def kmeans_lloyd(data: np.ndarray, k: int, max_iter: int = 10, init: np.ndarray | None = None):
    """Returns (centroids f32[K,D], labels i32[N], counts i32[K]).  ``k >= N`` follows gpu_ops.py:30-31."""
    from . import lib, _p
    n, d = data.shape
    if k >= n:
        return data.copy(), np.arange(n, dtype=np.int32), np.ones(n, dtype=np.int32)
    X = np.ascontiguousarray(data.astype(np.float32))
    C = np.ascontiguousarray((kmeans_reference_init(X, k) if init is None else init).astype(np.float32).copy())
    labels = np.zeros(n, dtype=np.int32)
    counts = np.zeros(k, dtype=np.int32)
    lib().orc_kmeans_lloyd(_p(X, ctypes.c_float), _p(C, ctypes.c_float), _p(labels, ctypes.c_int32),
                           _p(counts, ctypes.c_int32), n, k, d, max_iter)
    return C, labels, counts
