"""GPU parity: SOR with the reference's CPU-path semantics (scipy cKDTree, float64, data_processor.py:155-180)
against the oracle (which calls SciPy itself) and the golden fixtures -- bit-exact float32 mean distances."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("kind", ["mixed", "uniform", "clustered"])
@pytest.mark.parametrize("n,k", [(100_000, 16), (100_000, 27), (20_000, 50), (20_000, 5), (300_000, 8)])
def test_ckdtree_semantics_match_scipy(kind, n, k, cuda, gsx_lib):
    import torch
    import oracle
    from gsx import sor, synth
    xyz = synth.xyz(n, kind)
    want = oracle.sor_ckdtree_mean_dists(xyz, k)
    mask, means = sor.ckdtree_filter(torch.from_numpy(xyz).to(cuda), k, 2.0, return_means=True)
    got = means.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.flatnonzero(got != want)[:5]
    assert np.array_equal(mask.cpu().numpy(), oracle.threshold_mask(want, 2.0))


def test_ckdtree_goldens_and_host_entry(cuda, gsx_lib):
    from gsx import sor, synth
    g1 = np.load(G / "g1_100k.npz")
    xyz = synth.xyz(100_000, "mixed")
    for k, sigmas in ((16, (1.0, 2.0, 3.0)), (27, (20.0 - 4 * (17.0 / 9),))):
        for s in sigmas:
            mask, means = sor.ckdtree_filter_host(xyz, k, s, return_means=True)
            assert sha(means) == str(g1[f"ckd_k{k}_sha"])
            assert np.array_equal(np.packbits(mask), g1[f"ckd_k{k}_s{s:.3f}_mask"])


def test_ckdtree_edge_cases(cuda, gsx_lib):
    import torch
    import oracle
    from gsx import sor
    rng = np.random.default_rng(9)
    cases = {
        "duplicates": np.repeat(rng.normal(size=(400, 3)), 5, axis=0),
        "planar": np.c_[rng.uniform(-1, 1, (4000, 2)), np.zeros(4000)],
        "two_far_blobs": np.r_[rng.normal(0, 0.01, (2000, 3)), rng.normal(500, 0.01, (2000, 3))],
        "fewer_than_k": rng.normal(size=(10, 3)),       # cKDTree pads with inf -> mean inf
        "line": np.c_[np.linspace(0, 1, 3000), np.zeros(3000), np.zeros(3000)],
    }
    for name, pts in cases.items():
        xyz = pts.astype(np.float32)
        for k in (3, 16):
            want = oracle.sor_ckdtree_mean_dists(xyz, k)
            got = sor.ckdtree_mean_dists(torch.from_numpy(xyz).to(cuda), k).cpu().numpy()
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, k)


def test_ckdtree_1m(cuda, gsx_lib):
    import torch
    import oracle
    from gsx import sor, synth
    xyz = synth.xyz(1_000_000, "mixed")
    want = oracle.sor_ckdtree_mean_dists(xyz, 16)
    mask, means = sor.ckdtree_filter(torch.from_numpy(xyz).to(cuda), 16, 2.0, return_means=True)
    assert np.array_equal(means.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert int((~mask).sum()) == 2147   # SURVEY §8(c) anchor (cKDTree, k=16, sigma=2, 1 M mixed)
