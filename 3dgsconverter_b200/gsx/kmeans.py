"""Lloyd K-Means on device buffers: host plumbing over gsx_kmeans_* (gpu_ops.py:27-46,178-191)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import os

from ._abi import lib, check
from .sor import _ptr, _stream


def set_prefilter(on: bool):
    """Exact fma pre-filter of the assign step (labels bit-identical either way); see csrc/gsx_kmeans.cu."""
    lib.gsx_kmeans_set_prefilter(1 if on else 0)


def prefilter_enabled() -> bool:
    return bool(lib.gsx_kmeans_get_prefilter())


if "GSX_KMEANS_PREFILTER" in os.environ:
    set_prefilter(os.environ["GSX_KMEANS_PREFILTER"] == "1")


def kmeans_lloyd_batched(X: torch.Tensor, row_off, K: int, max_iter: int, init: torch.Tensor):
    """`nprob` independent problems stored back to back in X[*,D] (rows row_off[p]:row_off[p+1]),
    each with K centroids.  init: float32 [nprob,K,D] (consumed as the start, not modified).
    Returns (C [nprob,K,D], labels int32 [N] (problem-local ids), counts int32 [nprob,K])."""
    if not X.is_cuda or X.dtype != torch.float32 or not X.is_contiguous() or X.dim() != 2:
        raise ValueError("X must be a contiguous float32 CUDA tensor [N,D]")
    row_off = np.ascontiguousarray(row_off, dtype=np.int64)
    nprob = len(row_off) - 1
    D = X.shape[1]
    Cc = init.to(device=X.device, dtype=torch.float32).reshape(nprob, K, D).contiguous().clone()
    labels = torch.zeros(X.shape[0], dtype=torch.int32, device=X.device)
    counts = torch.zeros(nprob * K, dtype=torch.int32, device=X.device)
    ws = torch.empty(lib.gsx_kmeans_workspace_bytes(X.shape[0], nprob, K, D), dtype=torch.uint8, device=X.device)
    check(lib.gsx_kmeans_lloyd_device(_ptr(X), row_off.ctypes.data_as(C.POINTER(C.c_int64)), nprob, K, D, max_iter,
                                      _ptr(Cc), _ptr(labels), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
          "gsx_kmeans_lloyd_device")
    return Cc, labels, counts.reshape(nprob, K)


def kmeans_lloyd(X: torch.Tensor, K: int, max_iter: int, init: torch.Tensor):
    """Single problem.  Returns (C [K,D], labels [N], counts [K])."""
    Cc, labels, counts = kmeans_lloyd_batched(X, [0, X.shape[0]], K, max_iter, init.reshape(1, K, -1))
    return Cc[0], labels, counts[0]


def kmeans_host(data: np.ndarray, K: int, max_iter: int, init: np.ndarray):
    """Host-buffer entry (copies inside libgsx): binding target for gpu_ops.kmeans' GPU path."""
    X = np.ascontiguousarray(data, dtype=np.float32)
    n, D = X.shape
    Cc = np.ascontiguousarray(init, dtype=np.float32).copy()
    labels = np.zeros(n, dtype=np.int32)
    check(lib.gsx_kmeans_host(X.ctypes.data_as(C.c_void_p), n, K, D, max_iter, Cc.ctypes.data_as(C.c_void_p),
                              labels.ctypes.data_as(C.c_void_p)), "gsx_kmeans_host")
    return Cc, labels
