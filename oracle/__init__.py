"""CPU oracle for the gsx hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.
The product (``3dgsconverter_b200/``) never imports it.

Parity status.  The reference ships no tests or golden vectors and the real taichi wheel
cannot be installed here, so nothing below was ever compared with a Taichi RUN.  What is pinned, and to what:
  * Taichi-semantics SOR and Lloyd K-Means: to outputs of the reference's OWN kernel source --
    /root/reference/gsconverter/processing/gpu_ops.py executed unmodified under a serial stand-in for the
    `taichi` module (tests/golden/ti_serial.py, assumptions T1-T5 stated there: i32/f32 defaults, wrapping
    i32 products, strict IEEE per operation, serial atomics) -> tests/golden/g4_reference_kernels.npz
    (11 SOR runs on 8 clouds <= 3000 points, 4 K-Means problems) and g5_reference_sor_100k.npz (BASELINE
    configs[0]: the 100 k cloud, k=27 / --sor_intensity 5 and k=16) and g6_reference_kmeans_c3shape.npz (one
    20 011 x 45, K=256 Lloyd problem -- the C3 / tensor-core shape), tests/test_reference_kernels_pin.py.
    g7_reference_sor_1m_sample.npz: 30 000 queries of the 1 M cloud (the regime where int32-wrap != int64);
    g8_reference_sor_10m_sample.npz: 50 000 queries of the 10 M cloud of BASELINE configs[1].
    Beyond these fixtures the parity rests on this restatement plus the survey's anchor counts ("parity unpinned").
  * density / alpha / bbox / cKDTree SOR arithmetic: to the imported reference (tests/golden/make_goldens.py);
  * NumPy-visible arithmetic (pairwise mean/std, promotion rules): to NumPy itself (tests/test_oracle_pins.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "_build" / "liborc.so"
    src = _HERE / "gsx_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "_build/liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.orc_sor_mean_dists.argtypes = [f32p, i32p, i32p, f32p, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int,
                                         i64p]
        L.orc_sor_mean_dists.restype = None
        L.orc_pairwise_sum_f32.argtypes = [f32p, ctypes.c_int64]
        L.orc_pairwise_sum_f32.restype = ctypes.c_float
        L.orc_mean_std_f32.argtypes = [f32p, ctypes.c_int64, f32p]
        L.orc_mean_std_f32.restype = None
        L.orc_pairwise_mean_f64_rows.argtypes = [f64p, ctypes.c_int64, ctypes.c_int64, f32p]
        L.orc_pairwise_mean_f64_rows.restype = None
        L.orc_kmeans_assign.argtypes = [f32p, f32p, i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
        L.orc_kmeans_assign.restype = None
        L.orc_kmeans_update.argtypes = [f32p, f32p, i32p, i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
        L.orc_kmeans_update.restype = None
        L.orc_kmeans_lloyd.argtypes = [f32p, f32p, i32p, i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int32]
        L.orc_kmeans_lloyd.restype = None
        L.orc_powf.argtypes = [ctypes.c_float, ctypes.c_float]
        L.orc_powf.restype = ctypes.c_float
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def set_threads(n: int | None = None):
    if n:
        os.environ["OMP_NUM_THREADS"] = str(n)


from .sor import (sor_cell_size, sor_taichi_mean_dists, sor_taichi_mask, sor_ckdtree_mean_dists,  # noqa: E402
                  sor_ckdtree_mask, mean_std_f32, sor_slider, threshold_mask)
from .kmeans import kmeans_lloyd, kmeans_reference_init  # noqa: E402
from .filters import density_mask, density_slider, alpha_mask, bbox_mask  # noqa: E402
