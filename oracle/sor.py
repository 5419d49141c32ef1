"""Oracle: Statistical Outlier Removal, both reference semantics (TEST INFRASTRUCTURE).

``sor_taichi_*`` restates ``filter_sor_gpu`` (/root/reference/gsconverter/processing/
gpu_ops.py:193-263) around the C restatement of the Taichi kernel (:98-176);
``sor_ckdtree_*`` restates the CPU fallback of ``DataProcessor.remove_flyers``
(data_processor.py:155-180) -- keeping the mask the reference computes and then
discards.  NumPy-2 promotion semantics are the contract.
"""
from __future__ import annotations

import ctypes
import numpy as np

_P1, _P2, _P3 = 73856093, 19349663, 83492791


def sor_slider(intensity: float):
    """data_processor.py:125-134: --sor_intensity -> (k, sigma)."""
    k = int(10 + (intensity - 1) * (40 / 9))
    sigma = 20.0 - (intensity - 1) * (17.0 / 9)
    return k, sigma


def sor_cell_size(pos: np.ndarray):
    """gpu_ops.py:203-213.  Returns (min_bound f32[3], cell_size python float)."""
    n = pos.shape[0]
    lo = np.min(pos, axis=0)
    hi = np.max(pos, axis=0)
    vol = np.prod(hi - lo)
    if vol <= 0:
        vol = 1.0
    avg = max(1e-8, vol / n)
    cell = float((avg * 32) ** (1.0 / 3.0))
    return lo, max(cell, 1e-4)


def sor_hash_table(pos: np.ndarray, lo: np.ndarray, cell: float):
    """gpu_ops.py:216-237: int64 host hash, sort, bucket start/count tables."""
    n = pos.shape[0]
    gi = np.floor((pos - lo) / cell).astype(np.int32).astype(np.int64)
    hashed = (((gi[:, 0] * _P1) ^ (gi[:, 1] * _P2) ^ (gi[:, 2] * _P3)) % n).astype(np.int32)
    order = np.argsort(hashed, kind="stable")
    sh = hashed[order]
    uniq, first, cnt = np.unique(sh, return_index=True, return_counts=True)
    cell_start = np.full(n, -1, dtype=np.int32)
    cell_count = np.zeros(n, dtype=np.int32)
    cell_start[uniq] = first
    cell_count[uniq] = cnt
    return order, cell_start, cell_count


def sor_taichi_mean_dists(data: np.ndarray, k: int, hash_mode: str = "i32wrap", want_visits: bool = False):
    """final_means float32[N] in the caller's point order (gpu_ops.py:193-256)."""
    from . import lib, _p
    if data.ndim != 2 or data.shape[1] != 3:
        raise ValueError("Requires 3D data")
    pos = np.ascontiguousarray(data.astype(np.float32))
    n = pos.shape[0]
    lo, cell = sor_cell_size(pos)
    order, cell_start, cell_count = sor_hash_table(pos, lo, cell)
    spos = np.ascontiguousarray(pos[order])
    md = np.zeros(n, dtype=np.float32)
    visits = np.zeros(n, dtype=np.int64) if want_visits else None
    kk = min(int(k), 50)
    mode = {"i32wrap": 0, "i64": 1}[hash_mode]
    lib().orc_sor_mean_dists(_p(spos, ctypes.c_float), _p(cell_start, ctypes.c_int32), _p(cell_count, ctypes.c_int32),
                             _p(md, ctypes.c_float), float(lo[0]), float(lo[1]), float(lo[2]),
                             ctypes.c_float(cell), n, n, kk, mode,
                             _p(visits, ctypes.c_int64) if want_visits else None)
    final = np.zeros(n, dtype=np.float32)
    final[order] = md
    if want_visits:
        v = np.zeros(n, dtype=np.int64)
        v[order] = visits
        return final, v
    return final


def mean_std_f32(a: np.ndarray):
    """C restatement of np.mean/np.std on a float32 vector (A.1 step 9)."""
    from . import lib, _p
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.zeros(2, dtype=np.float32)
    lib().orc_mean_std_f32(_p(a, ctypes.c_float), a.shape[0], _p(out, ctypes.c_float))
    return out[0], out[1]


def threshold_mask(means: np.ndarray, threshold_factor: float):
    """gpu_ops.py:259-263 / data_processor.py:176-180, evaluated by NumPy itself."""
    gm = np.mean(means)
    gs = np.std(means)
    thresh = gm + threshold_factor * gs
    return means < thresh


def sor_taichi_mask(data, k=25, threshold_factor=1.0, hash_mode="i32wrap"):
    return threshold_mask(sor_taichi_mean_dists(data, k, hash_mode), threshold_factor)


def sor_ckdtree_mean_dists(coords: np.ndarray, k: int, chunk: int = 50000, workers: int = -1):
    """data_processor.py:160-173: exact (k+1)-NN in float64, mean of neighbours 1..k -> float32."""
    from scipy.spatial import cKDTree
    tree = cKDTree(coords)
    n = coords.shape[0]
    out = np.zeros(n, dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        d, _ = tree.query(coords[s:e], k=k + 1, workers=workers)
        out[s:e] = np.mean(d[:, 1:], axis=1)
    return out


def sor_ckdtree_mask(coords, k=25, threshold_factor=10.5):
    return threshold_mask(sor_ckdtree_mean_dists(coords, k), threshold_factor)
