"""CPU: pin the oracle (and libgsx's host helpers) against NumPy itself.

The reference has no tests; what can be pinned without Taichi is every piece of
NumPy-visible arithmetic on the path: float32 pairwise mean/std (gpu_ops.py:259-260), the float64
row mean (data_processor.py:172), the NumPy-2 promotion rules and the cell-size
expression (gpu_ops.py:205-213)."""
import ctypes as C

import numpy as np
import pytest

import oracle


@pytest.mark.parametrize("n", [1, 2, 5, 7, 8, 9, 15, 16, 27, 100, 127, 128, 129, 136, 255, 256, 257, 1000, 4097,
                               8191, 100_003, 1_000_000, 3_000_001, 16_777_216 + 5, 20_000_003])
def test_pairwise_mean_std_bit_exact(n):
    rng = np.random.default_rng(n)
    a = rng.gamma(2.0, 0.3, n).astype(np.float32)
    m, s = oracle.mean_std_f32(a)
    assert m.view(np.uint32) == np.mean(a).view(np.uint32)
    assert s.view(np.uint32) == np.std(a).view(np.uint32)
    b = (rng.standard_normal(n) * 1e3).astype(np.float32)
    m, s = oracle.mean_std_f32(b)
    assert m.view(np.uint32) == np.mean(b).view(np.uint32)
    assert s.view(np.uint32) == np.std(b).view(np.uint32)


def test_pairwise_random_lengths():
    rng = np.random.default_rng(0)
    for n in rng.integers(1, 200_000, 60):
        a = rng.random(int(n), dtype=np.float32)
        m, s = oracle.mean_std_f32(a)
        assert (m.view(np.uint32), s.view(np.uint32)) == (np.mean(a).view(np.uint32), np.std(a).view(np.uint32)), n


@pytest.mark.parametrize("k", [5, 10, 16, 25, 27, 50])
def test_row_mean_f64(k):
    rng = np.random.default_rng(k)
    d = np.sort(rng.random((2000, k + 1)), axis=1)
    out = np.zeros(2000, np.float32)
    oracle.lib().orc_pairwise_mean_f64_rows(d.ctypes.data_as(C.POINTER(C.c_double)), 2000, k + 1,
                                            out.ctypes.data_as(C.POINTER(C.c_float)))
    want = np.zeros(2000, np.float32)
    want[:] = np.mean(d[:, 1:], axis=1)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_numpy2_promotion_facts():
    """the contract is NumPy-2 (NEP 50) semantics."""
    assert int(np.__version__.split(".")[0]) >= 2
    a = np.array([0.1], np.float32)
    assert (a <= 0.1)[0]                      # python float is weak -> compared in float32
    assert not (a <= np.float64(0.1))[0]      # np.float64 is strong -> compared in float64
    assert type(np.float32(1) + 2.5 * np.float32(1)) is np.float32
    assert type((np.float32(3) * 32) ** (1.0 / 3.0)) is np.float32
    assert (a / 1.1).dtype == np.float32
    assert type(np.float32(7) / 3) is np.float32


def _numpy_cell(lo, hi, n):
    extent = hi - lo
    vol = np.prod(extent)
    if vol <= 0:
        vol = 1.0
    avg = max(1e-8, vol / n)
    cell = float((avg * 32) ** (1.0 / 3.0))
    return max(cell, 1e-4)


def test_cell_size_helper_matches_numpy(gsx_lib):
    rng = np.random.default_rng(7)
    for t in range(3000):
        scale = 10.0 ** rng.uniform(-6, 4)
        lo = (rng.standard_normal(3) * scale).astype(np.float32)
        hi = lo + (rng.random(3) * scale).astype(np.float32) * (rng.random() < 0.97)
        hi = hi.astype(np.float32)
        n = int(10 ** rng.uniform(0, 9.3))
        mm = np.r_[lo, hi].astype(np.float32)
        got = gsx_lib.gsx_sor_cell_size(mm.ctypes.data_as(C.POINTER(C.c_float)), n)
        want = np.float32(_numpy_cell(lo, hi, n))
        assert np.float32(got) == want, (lo, hi, n, got, want)
    # the oracle's driver goes through NumPy; spot-check it agrees on a real cloud
    pts = rng.standard_normal((1000, 3)).astype(np.float32)
    lo, cell = oracle.sor_cell_size(pts)
    mm = np.r_[pts.min(0), pts.max(0)].astype(np.float32)
    assert np.float32(gsx_lib.gsx_sor_cell_size(mm.ctypes.data_as(C.POINTER(C.c_float)), 1000)) == np.float32(cell)


def test_alpha_threshold_helper_matches_numpy(gsx_lib):
    for m in list(range(1, 255)) + [0.5, 254.9, 1e-9]:
        a = np.clip(m / 255.0, 1e-6, 1.0 - 1e-6)
        want = float(np.log(a / (1.0 - a)))
        got = gsx_lib.gsx_alpha_logit_threshold(float(m))
        # libm vs NumPy's SIMD log: at most one float64 ulp apart, i.e. the same float32 cut
        assert abs(got - want) <= np.spacing(abs(want)), m
        assert np.float32(got) == np.float32(want)


def test_sor_slider_and_k_cap():
    assert oracle.sor_slider(5) == (27, 20.0 - 4 * (17.0 / 9))   # k=27, not 25
    assert oracle.sor_slider(1) == (10, 20.0)
    assert oracle.sor_slider(10)[0] == 50
    pts = np.random.default_rng(1).random((300, 3)).astype(np.float32)
    a = oracle.sor_taichi_mean_dists(pts, 50)
    b = oracle.sor_taichi_mean_dists(pts, 80)  # K capped at 50 (gpu_ops.py:244)
    assert np.array_equal(a, b)
