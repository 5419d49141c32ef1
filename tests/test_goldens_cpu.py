"""CPU: the oracle reproduces the committed golden vectors (tests/golden/*.npz, made by
make_goldens.py in the container that has /root/reference)."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

import oracle
from gsx import synth

G = Path(__file__).resolve().parent / "golden"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g1():
    return np.load(G / "g1_100k.npz")


@pytest.fixture(scope="module")
def cloud():
    return synth.xyz(100_000, "mixed"), synth.attributes(100_000)


def test_synth_is_pinned(g1, cloud):
    assert sha(cloud[0]) == str(g1["xyz_sha"]) and sha(cloud[1]["opacity"]) == str(g1["opacity_sha"])


def test_density_alpha_bbox_goldens(g1, cloud):
    xyz, at = cloud
    for sens in (0.1, 0.5, 0.9):
        for multi in (False, True):
            m, _ = oracle.density_mask(xyz, sensitivity=sens, keep_multicluster=multi)
            assert np.array_equal(np.packbits(m), g1[f"density_s{sens}_m{int(multi)}"])
    m, _ = oracle.density_mask(xyz, 0.7, 0.05, None, True)
    assert np.array_equal(np.packbits(m), g1["density_v0.7_t0.05_m1"])
    for mo in (1, 5, 128):
        assert np.array_equal(np.packbits(oracle.alpha_mask(at["opacity"], mo)), g1[f"alpha_{mo}"])
    assert np.array_equal(np.packbits(oracle.bbox_mask(xyz[:, 0], xyz[:, 1], xyz[:, 2], -2, -2, -2, 2, 2, 2)),
                          g1["bbox_2"])


@pytest.mark.parametrize("k", [16, 27])
def test_sor_goldens_both_semantics(g1, cloud, k):
    xyz, _ = cloud
    tai = oracle.sor_taichi_mean_dists(xyz, k, "i32wrap")
    assert sha(tai) == str(g1[f"tai_k{k}_sha"])
    ckd = oracle.sor_ckdtree_mean_dists(xyz, k)
    assert sha(ckd) == str(g1[f"ckd_k{k}_sha"])
    sigmas = (1.0, 2.0, 3.0) if k == 16 else (oracle.sor_slider(5)[1],)
    for s in sigmas:
        assert np.array_equal(np.packbits(oracle.threshold_mask(tai, s)), g1[f"tai_k{k}_s{s:.3f}_mask"])
        assert np.array_equal(np.packbits(oracle.threshold_mask(ckd, s)), g1[f"ckd_k{k}_s{s:.3f}_mask"])
    if k == 16:
        assert np.array_equal(tai, g1["tai_k16_means"]) and np.array_equal(ckd, g1["ckd_k16_means"])


def test_kmeans_goldens():
    g2 = np.load(G / "g2_kmeans.npz")
    X = synth.attributes(100_000)["f_rest"]
    assert sha(X) == str(g2["sh45_k16_it10_X_sha"])
    np.random.seed(1234)
    C, L, cnt = oracle.kmeans_lloyd(X, 16, 10)  # consumes exactly one np.random.choice draw (gpu_ops.py:182)
    assert np.array_equal(C, g2["sh45_k16_it10_C"]) and np.array_equal(cnt, g2["sh45_k16_it10_counts"])
    assert sha(L) == str(g2["sh45_k16_it10_labels_sha"])
    X1 = synth.attributes(50_000)["scale"].reshape(-1, 1)[:50_000].copy()
    C, L, cnt = oracle.kmeans_lloyd(X1, 256, 20, init=g2["scale1_k256_it20_init"])
    assert np.array_equal(C, g2["scale1_k256_it20_C"]) and sha(L) == str(g2["scale1_k256_it20_labels_sha"])


def test_kmeans_k_ge_n_passthrough():
    X = np.arange(12, dtype=np.float32).reshape(4, 3)
    C, L, _ = oracle.kmeans_lloyd(X, 4, 10)
    assert np.array_equal(C, X) and np.array_equal(L, np.arange(4, dtype=np.int32))
