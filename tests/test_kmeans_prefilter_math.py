"""CPU: the rounding-error margin of the K-Means pre-filter (csrc/gsx_kmeans.cu, k_kmeans_assign_pre) always keeps
the contract's answer in the candidate set.  The kernel's score chain (one fma per dim, float32) and its margin
formula are restated here in NumPy (fma emulated exactly through float64); the contract's label comes from the
oracle.  Adversarial inputs: near-duplicate centroids, large common offsets (cancellation), tiny and huge scales."""
import numpy as np
import pytest

import oracle

f32 = np.float32
U = f32(5.9604645e-8)


def fma32(a, b, c):
    """float32 fma for float32 arrays: a*b is exact in float64, one rounding to float32 at the end."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def kernel_scores_and_margin(X, C):
    n, D = X.shape
    K = C.shape[0]
    cn = np.zeros(K, f32)
    for d in range(D):
        cn = fma32(C[:, d], C[:, d], cn)
    half = (f32(-0.5) * cn).astype(f32)
    xn = np.zeros(n, f32)
    for d in range(D):
        xn = fma32(X[:, d], X[:, d], xn)
    s = np.broadcast_to(half, (n, K)).copy()
    for d in range(D):
        s = fma32(np.broadcast_to(X[:, d:d + 1], (n, K)), np.broadcast_to(C[:, d], (n, K)), s)
    # margin exactly as in k_kmeans_assign_pre
    Cm = f32(np.sqrt(cn.max(), dtype=np.float32) * f32(1.0001))
    xnu = (xn * f32(1.0001)).astype(f32)
    xnorm = (np.sqrt(xnu, dtype=np.float32) * f32(1.0001)).astype(f32)
    kGam2D = f32((2 * D + 2) * U * f32(1.02))
    kGs = f32(2.0 * (D + 3) * U * f32(1.02))
    delta = (kGam2D * (Cm * Cm + f32(2) * xnorm * Cm) * f32(1.01) + f32(1e-37)).astype(f32)
    smax = s.max(axis=1)
    e_ub = np.maximum(xnu - f32(2) * smax + delta, f32(0)).astype(f32)
    marg = (f32(2) * delta + kGs * e_ub + f32(1e-37)).astype(f32)
    return s, smax, marg


def _cases():
    rng = np.random.default_rng(0)
    proto = rng.normal(0, 0.15, (64, 45)).astype(f32)
    X = (proto[rng.integers(0, 64, 4000)] + rng.normal(0, 0.03, (4000, 45))).astype(f32)
    yield "sh_like", X, X[rng.choice(4000, 256, replace=False)]
    C = X[rng.choice(4000, 128, replace=False)].copy()
    C2 = np.r_[C, C + f32(1e-6) * rng.standard_normal(C.shape).astype(f32)]       # near-duplicate centroids
    yield "near_duplicate_centroids", X, C2.astype(f32)
    off = f32(100.0)
    yield "large_offset_cancellation", (X + off).astype(f32), (C + off).astype(f32)
    yield "tiny_scale", (X * f32(1e-12)).astype(f32), (C * f32(1e-12)).astype(f32)
    yield "huge_scale", (X * f32(3e10)).astype(f32), (C * f32(3e10)).astype(f32)   # d^2 ~ 1e19..1e21: around 1e20
    Xg = rng.integers(-3, 4, (3000, 9)).astype(f32)                               # lattice: exact ties everywhere
    yield "lattice_ties_d9", Xg, Xg[rng.choice(3000, 100, replace=False)]
    X24 = rng.standard_normal((3000, 24)).astype(f32)
    yield "gauss_d24", X24, X24[rng.choice(3000, 300, replace=False)]


@pytest.mark.parametrize("name,X,C", list(_cases()), ids=[c[0] for c in _cases()])
def test_margin_contains_contract_label(name, X, C):
    X = np.ascontiguousarray(X)
    C = np.ascontiguousarray(C)
    n, D = X.shape
    K = C.shape[0]
    labels = np.zeros(n, np.int32)
    import ctypes
    from oracle import _p
    oracle.lib().orc_kmeans_assign(_p(X, ctypes.c_float), _p(C, ctypes.c_float), _p(labels, ctypes.c_int32), n, K, D)
    s, smax, marg = kernel_scores_and_margin(X, C)
    ok = labels >= 0          # label -1: every strict distance is >= the 1e20 start value (gpu_ops.py:61); the
    rows = np.flatnonzero(ok)  # kernel returns -1 as well because no candidate can beat 1e20 either
    s_star = s[rows, labels[rows]]
    assert np.all(s_star >= (smax - marg)[rows]), (name, float(((smax - marg)[rows] - s_star).max()))
    if name == "huge_scale":
        assert (~ok).sum() > 0 and ok.sum() > 0      # the case straddles the 1e20 quirk on purpose
    ncand = (s >= (smax - marg)[:, None]).sum(axis=1)
    # the margin is tight enough to be useful on well-separated data (otherwise the pre-filter degenerates)
    if name in ("sh_like", "gauss_d24"):
        assert ncand.mean() < 1.5, ncand.mean()
