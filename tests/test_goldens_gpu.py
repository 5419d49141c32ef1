"""GPU: the CUDA path against the committed golden fixtures (no oracle needed at run time), at
BASELINE.json's full single-GPU size too (10 M splats: digests made once by the CPU oracle)."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_g1_100k_all_filters(cuda, gsx_lib):
    import torch
    from gsx import sor, density, masks, synth
    g1 = np.load(G / "g1_100k.npz")
    xyz = synth.xyz(100_000, "mixed")
    op = synth.attributes(100_000)["opacity"]
    x = torch.from_numpy(xyz).to(cuda)
    for k, sigmas in ((16, (1.0, 2.0, 3.0)), (27, (20.0 - 4 * (17.0 / 9),))):
        for s in sigmas:
            for mode in ("i32wrap", "i64"):
                mask, means = sor.sor_filter(x, k, s, hash_mode=mode, return_means=True)
                assert sha(means.cpu().numpy()) == str(g1[f"tai_k{k}_sha"])
                assert np.array_equal(np.packbits(mask.cpu().numpy()), g1[f"tai_k{k}_s{s:.3f}_mask"])
    for sens in (0.1, 0.5, 0.9):
        for multi in (False, True):
            m, _ = density.density_filter(x, sensitivity=sens, keep_multicluster=multi)
            assert np.array_equal(np.packbits(m.cpu().numpy()), g1[f"density_s{sens}_m{int(multi)}"])
    m, _ = density.density_filter(x, 0.7, 0.05, None, True)
    assert np.array_equal(np.packbits(m.cpu().numpy()), g1["density_v0.7_t0.05_m1"])
    o = torch.from_numpy(op).to(cuda)
    for mo in (1, 5, 128):
        assert np.array_equal(np.packbits(masks.alpha_mask(o, mo).cpu().numpy()), g1[f"alpha_{mo}"])
    assert np.array_equal(np.packbits(masks.bbox_mask(x, -2, -2, -2, 2, 2, 2).cpu().numpy()), g1["bbox_2"])


def test_g1_1m_digests(cuda, gsx_lib):
    import torch
    from gsx import sor, synth
    g = np.load(G / "g1_1m.npz")
    xyz = synth.xyz(1_000_000, "mixed")
    assert sha(xyz) == str(g["xyz_sha"])
    x = torch.from_numpy(xyz).to(cuda)
    for mode in ("i32wrap", "i64"):
        for s in (2.0, 3.0):
            mask, means = sor.sor_filter(x, 16, s, hash_mode=mode, return_means=True)
            assert sha(means.cpu().numpy()) == str(g[f"tai_k16_{mode}_sha"])
            assert sha(np.packbits(mask.cpu().numpy())) == str(g[f"tai_k16_{mode}_s{s:.1f}_mask_sha"])


def test_g2_kmeans(cuda, gsx_lib):
    import torch
    from gsx import kmeans as gk, synth
    g2 = np.load(G / "g2_kmeans.npz")
    X = synth.attributes(100_000)["f_rest"]
    for name, k, it in (("sh45_k16_it10", 16, 10), ("sh45_k256_it10", 256, 10)):
        C, L, cnt = gk.kmeans_lloyd(torch.from_numpy(X).to(cuda), k, it, torch.from_numpy(g2[f"{name}_init"]).to(cuda))
        # contract (north_star): centroids within 1e-5 relative; we are bit-identical
        assert np.allclose(C.cpu().numpy(), g2[f"{name}_C"], rtol=1e-5, atol=0)
        assert np.array_equal(C.cpu().numpy(), g2[f"{name}_C"])
        assert np.array_equal(cnt.cpu().numpy(), g2[f"{name}_counts"])
        assert sha(L.cpu().numpy()) == str(g2[f"{name}_labels_sha"])
    X1 = synth.attributes(50_000)["scale"].reshape(-1, 1)[:50_000].copy()
    C, L, cnt = gk.kmeans_lloyd(torch.from_numpy(X1).to(cuda), 256, 20,
                                torch.from_numpy(g2["scale1_k256_it20_init"]).to(cuda))
    assert np.array_equal(C.cpu().numpy(), g2["scale1_k256_it20_C"])


def test_g3_10m_full_size(cuda, gsx_lib):
    """BASELINE configs[1]: 10 M splats, SOR k=16 + density 0.5 (+ chained, converter.py:205-234 order)."""
    import torch
    from gsx import sor, density, synth
    p = G / "g3_10m.npz"
    if not p.exists():
        pytest.skip("g3_10m.npz not generated yet")
    g = np.load(p)
    xyz = synth.xyz(10_000_000, "mixed")
    assert sha(xyz) == str(g["xyz_sha"])
    x = torch.from_numpy(xyz).to(cuda)
    for mode in ("i32wrap", "i64"):
        if f"tai_k16_{mode}_sha" not in g:
            continue
        (mask, means), st = sor.sor_filter(x, 16, 2.0, hash_mode=mode, return_means=True), None
        m = means.cpu().numpy()
        assert np.array_equal(m[:1000], g[f"tai_k16_{mode}_head"]) and np.array_equal(m[-1000:], g[f"tai_k16_{mode}_tail"])
        assert sha(m) == str(g[f"tai_k16_{mode}_sha"])
        ms = sor.mean_std(means).cpu().numpy()
        assert np.array_equal(ms.view(np.uint32), g[f"tai_k16_{mode}_meanstd"].view(np.uint32))
        assert sha(np.packbits(mask.cpu().numpy())) == str(g[f"tai_k16_{mode}_s2.0_mask_sha"])
        assert int((~mask).sum()) == int(g[f"tai_k16_{mode}_s2.0_removed"])
    if "density_s0.5_m1_mask_sha" in g:
        dm, info = density.density_filter(x, sensitivity=0.5, keep_multicluster=True)
        assert sha(np.packbits(dm.cpu().numpy())) == str(g["density_s0.5_m1_mask_sha"])
        # chained: density survivors -> SOR (size-independent properties: idempotent masks, subset)
        xs = x[dm].contiguous()
        m2 = sor.sor_filter(xs, 16, 2.0)
        assert m2.shape[0] == int(dm.sum()) and 0 < int(m2.sum()) <= m2.shape[0]
        assert torch.equal(m2, sor.sor_filter(xs, 16, 2.0))  # run-to-run deterministic


def test_large_size_properties(cuda, gsx_lib):
    """Size-independent properties at a size the oracle cannot check quickly (30 M splats)."""
    import torch
    from gsx import sor
    g = torch.Generator(device=cuda).manual_seed(20260923)
    n = 30_000_000
    x = (torch.rand((n, 3), device=cuda, generator=g) * 20 - 10)
    x[: n // 3] = torch.randn((n // 3, 3), device=cuda, generator=g) * 0.3 + 2.0
    mask, means = sor.sor_filter(x, 16, 2.0, hash_mode="i64", return_means=True)
    assert torch.isfinite(means).all() and (means >= 0).all()
    ms = sor.mean_std(means)
    thresh = ms[0] + torch.tensor(2.0, device=cuda) * ms[1]
    assert torch.equal(mask, means < thresh)              # mask consistent with the means
    perm = torch.randperm(n, device=cuda, generator=g)
    mask_p, means_p = sor.sor_filter(x[perm].contiguous(), 16, 2.0, hash_mode="i64", return_means=True)
    assert torch.equal(means_p, means[perm])              # permutation equivariance of the mean distances
