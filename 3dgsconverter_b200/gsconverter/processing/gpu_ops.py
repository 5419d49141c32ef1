"""Drop-in for ``gsconverter/processing/gpu_ops.py`` backed by libgsx.so (hand-written sm_100a CUDA).

Same importable surface as the reference (gpu_ops.py:8-46,193-263):

    HAS_TAICHI                      bool   -- name kept because data_processor.py:144 and
                                              formats/sog.py:524 read it; True <=> the gsx CUDA backend
                                              is usable (a CUDA device is present)
    kmeans(data, k, max_iter=10, tolerance=1e-4, use_gpu=True, verbose=False) -> (centroids, labels)
    filter_sor_gpu(data_np, k=25, threshold_factor=1.0, verbose=False)        -> bool mask | None

Differences that are deliberate: there is NO silent CPU fallback.  libgsx.so missing => ImportError at
import; a failing CUDA call => GsxError propagates (the reference swallowed it and ran scikit-learn).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

try:
    import gsx  # noqa: F401
except ImportError:  # dropped into a foreign tree: locate the backend next to this repo or via GSX_HOME
    _home = os.environ.get("GSX_HOME") or str(Path(__file__).resolve().parents[2])
    if _home not in sys.path:
        sys.path.insert(0, _home)
    import gsx  # noqa: F401  (ImportError here is fatal on purpose: no CPU fallback)

from gsx import backend_available

HAS_TAICHI = backend_available()


def _hash_mode():
    # "i32wrap" reproduces the Taichi kernel's default-int arithmetic; "i64" is the intended hash.
    return os.environ.get("GSX_SOR_HASH", "i32wrap")


def kmeans(data: np.ndarray, k: int, max_iter=10, tolerance=1e-4, use_gpu=True, verbose=False):
    """gpu_ops.py:27-46.  ``tolerance`` is accepted and ignored, as in the reference (F9)."""
    n, d = data.shape
    if k >= n:  # gpu_ops.py:30-31
        return data.copy(), np.arange(n, dtype=np.int32)
    if not use_gpu:
        # explicit request for the reference's CPU algorithm (gpu_ops.py:34-38, 48-52): scikit-learn
        # MiniBatchKMeans, unseeded -- a different algorithm, outside the parity contract.
        from sklearn.cluster import MiniBatchKMeans
        km = MiniBatchKMeans(n_clusters=k, max_iter=max_iter, batch_size=min(4096 * 4, len(data)), n_init="auto",
                             compute_labels=True)
        km.fit(data)
        return km.cluster_centers_.astype(np.float32), km.labels_.astype(np.int32)
    if not HAS_TAICHI:
        raise RuntimeError("gsx: no CUDA device available and CPU fallback is disabled by design")
    from gsx import kmeans as _km
    x = data.astype(np.float32)
    # exactly the reference's draw from the global NumPy RNG (gpu_ops.py:182) so np.random.seed reproduces
    init = x[np.random.choice(n, k, replace=False)].astype(np.float32)
    centroids, labels = _km.kmeans_host(x, int(k), int(max_iter), init)
    return centroids, labels


def filter_sor_gpu(data_np: np.ndarray, k: int = 25, threshold_factor: float = 1.0, verbose=False):
    """gpu_ops.py:193-263: keep-mask of the statistical outlier removal, or None if no backend."""
    if not HAS_TAICHI:
        return None
    n, d = data_np.shape
    if d != 3:
        raise ValueError("Requires 3D data")
    from gsx import sor as _sor
    if os.environ.get("GSX_SOR_SEMANTICS", "taichi") == "ckdtree":
        # opt-in: the arithmetic of the reference's CPU path (data_processor.py:155-180, exact float64 KNN)
        return _sor.ckdtree_filter_host(data_np, int(k), float(threshold_factor))
    return _sor.sor_filter_host(data_np, int(k), float(threshold_factor), hash_mode=_hash_mode())
