"""Put the gsx backend behind an already-installed 3dgsconverter without touching its files.

    import gsx.dropin; gsx.dropin.patch()      # before running gsconverter.main / Converter

Replaces, in the *installed* ``gsconverter.processing`` package (converter.py:10,150 and
formats/sog.py:11 import from there):
  * ``gpu_ops.kmeans``, ``gpu_ops.filter_sor_gpu``, ``gpu_ops.HAS_TAICHI``
  * the ``DataProcessor`` class (same public surface; the filters run on libgsx and keep their
    working set in HBM; ``defer=True`` gathers the host records once, when ``.data`` is read)
Host-only helpers and everything else of the reference stay as they are.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent.parent


def _load_ours(modname: str, relpath: str):
    spec = importlib.util.spec_from_file_location(modname, _HERE / relpath,
                                                  submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    return spec, mod


class _GsxCodebookKMeans:
    """Opt-in stand-in for the scikit-learn MiniBatchKMeans that formats/sog.py:561 runs on the flattened shN
    centroids (~737 k scalars -> 256): exact Lloyd on the GPU (gsx 1-D kernel), init = the same kind of draw
    gpu_ops.kmeans makes.  A different algorithm than MiniBatchKMeans (which is itself unseeded), so it is not the
    default; `patch(codebook="gpu")` installs it."""

    def __init__(self, n_clusters=8, n_init="auto", max_iter=20, **_):
        self.n_clusters, self.max_iter = int(n_clusters), int(max_iter)

    def fit(self, X):
        import numpy as np
        from gsx import kmeans as _km
        x = np.ascontiguousarray(X, dtype=np.float32).reshape(len(X), -1)
        k = min(self.n_clusters, len(x))
        init = x[np.random.choice(len(x), k, replace=False)]
        C, L = _km.kmeans_host(x, k, self.max_iter, init)
        self.cluster_centers_, self.labels_ = C, L
        return self


def patch(verbose: bool = False, defer: bool = True, codebook: str = "sklearn", require_cuda: bool = True):
    """require_cuda: refuse (return False, leave the reference untouched) when no CUDA device is usable, so that a
    CPU-only host keeps the reference's own SciPy / scikit-learn paths (there is no CPU fallback inside gsx)."""
    if require_cuda:
        from . import backend_available
        if not backend_available():
            if verbose:
                print("[gsx] no CUDA device: gsconverter.processing left unpatched")
            return False
    ref_gpu_ops = importlib.import_module("gsconverter.processing.gpu_ops")
    ref_dp = importlib.import_module("gsconverter.processing.data_processor")
    if getattr(ref_gpu_ops, "_GSX_PATCHED", False):
        return True
    # load our two modules under private names but with the reference's package as parent, so their
    # relative imports (..utils.utility_functions) resolve to the installed reference
    ours = {}
    for name, rel in (("gpu_ops", "gsconverter/processing/gpu_ops.py"),
                      ("data_processor", "gsconverter/processing/data_processor.py")):
        full = f"gsconverter.processing._gsx_{name}"
        spec = importlib.util.spec_from_file_location(full, _HERE / rel)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "gsconverter.processing"
        sys.modules[full] = mod
        ours[name] = (spec, mod)
    ours["gpu_ops"][0].loader.exec_module(ours["gpu_ops"][1])
    g = ours["gpu_ops"][1]
    ref_gpu_ops.kmeans = g.kmeans
    ref_gpu_ops.filter_sor_gpu = g.filter_sor_gpu
    ref_gpu_ops.HAS_TAICHI = g.HAS_TAICHI
    ref_gpu_ops._GSX_PATCHED = True
    # our data_processor does `from .gpu_ops import ...`: that now resolves to the patched reference module
    ours["data_processor"][0].loader.exec_module(ours["data_processor"][1])
    Ours = ours["data_processor"][1].DataProcessor
    # the whole class is replaced: the device-resident working set needs the `data` property
    Ours.defer_compaction = bool(defer)   # converter.py ignores the filters' return values (converter.py:194-259)
    ref_dp.DataProcessor = Ours
    proc_pkg = importlib.import_module("gsconverter.processing")
    proc_pkg.DataProcessor = Ours
    conv = sys.modules.get("gsconverter.converter")
    if conv is not None and hasattr(conv, "DataProcessor"):
        conv.DataProcessor = Ours
    ref_gpu_ops._gsx_module = g     # batch-ahead statistics: gpu_ops._gsx_module.batch_stats
    sog = sys.modules.get("gsconverter.formats.sog")
    if sog is None:
        try:
            sog = importlib.import_module("gsconverter.formats.sog")
        except Exception:  # noqa: BLE001  (optional dependency of the writer missing: nothing to patch there)
            sog = None
    if sog is not None and codebook == "gpu" and hasattr(sog, "MiniBatchKMeans"):
        sog.MiniBatchKMeans = _GsxCodebookKMeans      # sog.py:561 -> exact 1-D Lloyd on the GPU
    if verbose:
        print("[gsx] gsconverter.processing patched: SOR / density / bbox / alpha / K-Means run on libgsx.so")
    return True
