"""GPU parity for K-Means (bit-exact vs the oracle's serial order), density, bbox and alpha masks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _kmeans_data(n, d):
    from gsx import synth
    at = synth.attributes(n)
    if d == 1:
        return np.ascontiguousarray(at["scale"].reshape(-1, 1)[:n])
    return np.ascontiguousarray(at["f_rest"][:, :d])


@pytest.mark.parametrize("n,d,k,it", [(20_000, 45, 64, 5), (50_000, 1, 256, 20), (10_000, 9, 16, 10),
                                      (10_000, 24, 100, 3), (3_000, 3, 7, 4), (5_000, 5, 33, 3), (4_000, 45, 300, 2), (6_000, 3, 2100, 2),
                                      (300_000, 45, 256, 2), (2_500, 9, 1024, 3)])
def test_kmeans_matches_oracle(n, d, k, it, cuda, gsx_lib):
    import torch
    import oracle
    from gsx import kmeans as gk
    X = _kmeans_data(n, d)
    np.random.seed(1234)
    init = oracle.kmeans_reference_init(X, k)
    Co, Lo, cnto = oracle.kmeans_lloyd(X, k, it, init=init)
    C, L, cnt = gk.kmeans_lloyd(torch.from_numpy(X).to(cuda), k, it, torch.from_numpy(init).to(cuda))
    assert np.array_equal(L.cpu().numpy(), Lo)
    assert np.array_equal(cnt.cpu().numpy(), cnto)
    # contract: 1e-5 relative (north_star); we are bit-identical to the oracle's serial float32 order
    assert np.array_equal(C.cpu().numpy().view(np.uint32), Co.view(np.uint32))
    Ch, Lh = gk.kmeans_host(X, k, it, init)
    assert np.array_equal(Lh, Lo) and np.allclose(Ch, Co, rtol=1e-5, atol=0)


def test_kmeans_batched_chunks(cuda, gsx_lib):
    """The SOG shN schedule: independent chunks in one launch (sog.py:527-549)."""
    import torch
    import oracle
    from gsx import kmeans as gk, synth
    X = synth.attributes(30_000)["f_rest"]
    offs = [0, 7_000, 16_001, 30_000]
    K = 32
    rng = np.random.default_rng(3)
    init = np.stack([X[offs[p]:offs[p + 1]][rng.choice(offs[p + 1] - offs[p], K, replace=False)] for p in range(3)])
    C, L, cnt = gk.kmeans_lloyd_batched(torch.from_numpy(X).to(cuda), offs, K, 4, torch.from_numpy(init).to(cuda))
    for p in range(3):
        Co, Lo, cnto = oracle.kmeans_lloyd(X[offs[p]:offs[p + 1]], K, 4, init=init[p])
        assert np.array_equal(L[offs[p]:offs[p + 1]].cpu().numpy(), Lo)
        assert np.array_equal(C[p].cpu().numpy().view(np.uint32), Co.view(np.uint32))


def test_kmeans_empty_cluster_collapses_to_zero(cuda, gsx_lib):
    import torch
    import oracle
    from gsx import kmeans as gk
    X = np.r_[np.zeros((50, 2)), np.ones((50, 2))].astype(np.float32)
    init = np.array([[0, 0], [1, 1], [50, 50]], np.float32)  # third centroid attracts nothing
    Co, Lo, cnto = oracle.kmeans_lloyd(X, 3, 2, init=init)
    C, L, cnt = gk.kmeans_lloyd(torch.from_numpy(X).to(cuda), 3, 2, torch.from_numpy(init).to(cuda))
    assert cnto[2] == 0 and np.all(Co[2] == 0)
    assert np.array_equal(C.cpu().numpy(), Co) and np.array_equal(L.cpu().numpy(), Lo)


@pytest.mark.parametrize("n", [1000, 100_000, 1_000_000])
@pytest.mark.parametrize("sens,multi", [(0.1, False), (0.5, True), (0.5, False), (0.9, True)])
def test_density_matches_oracle(n, sens, multi, cuda, gsx_lib):
    import torch
    import oracle
    from gsx import density, synth
    xyz = synth.xyz(n, "mixed")
    want, info_o = oracle.density_mask(xyz, sensitivity=sens, keep_multicluster=multi)
    got, info = density.density_filter(torch.from_numpy(xyz).to(cuda), sensitivity=sens, keep_multicluster=multi)
    assert info["clusters"] == info_o["clusters"] and info["max_len"] == info_o["max_len"]
    assert np.array_equal(got.cpu().numpy(), want)


def test_density_explicit_params_and_hash_path(cuda, gsx_lib):
    import torch
    import oracle
    from gsx import density
    rng = np.random.default_rng(2)
    # far-flung sparse cloud: voxel grid too large for the dense histogram -> hash-table path
    xyz = np.r_[rng.normal(0, 1, (20_000, 3)), rng.uniform(-5e4, 5e4, (2_000, 3))].astype(np.float32)
    for vs, thr, multi in ((0.5, 0.05, True), (0.1, 0.0, False), (2.0, 1.0, False)):
        want, _ = oracle.density_mask(xyz, voxel_size=vs, threshold_percentage=thr, keep_multicluster=multi)
        got, _ = density.density_filter(torch.from_numpy(xyz).to(cuda), vs, thr, keep_multicluster=multi)
        assert np.array_equal(got.cpu().numpy(), want), (vs, thr, multi)


def test_density_huge_extent_wide_keys(cuda, gsx_lib):
    """a distant flyer with a small voxel spans >= 2^21 voxels per axis -- the reference (np.unique on int64
    triples) handles any extent; the hash path switches to two-word keys instead of failing."""
    import torch
    import oracle
    from gsx import density
    rng = np.random.default_rng(5)
    core = rng.normal(0, 1, (30_000, 3))
    far_clump = rng.normal(0, 0.02, (3_000, 3)) + np.array([1e7, -1e7, 1e7])   # dense, 1e8 voxels away at voxel 0.1
    xyz = np.r_[core, far_clump, np.array([[1e7, 1e7, 1e7], [-3e6, 0, 0]])].astype(np.float32)
    for vs, thr, multi in ((0.1, 0.01, True), (0.1, 0.01, False), (0.5, 0.5, True), (0.1, 0.0, True)):
        want, info_o = oracle.density_mask(xyz, voxel_size=vs, threshold_percentage=thr, keep_multicluster=multi)
        got, info = density.density_filter(torch.from_numpy(xyz).to(cuda), vs, thr, keep_multicluster=multi)
        assert info["clusters"] == info_o["clusters"], (vs, thr, multi)
        assert np.array_equal(got.cpu().numpy(), want), (vs, thr, multi)


def test_density_member_mask_bitmap_and_hash_sets(cuda, gsx_lib):
    """The keep set is a bitmap over the kept voxels' bounding box when that box has <= 2^27 voxels and a hash set
    otherwise (21-bit keys here; two-word keys in the test above): same mask, and both equal the oracle.  Two dense
    blobs 900 voxels apart on every axis (box 900^3 > 2^27 -> hash set when both are kept, bitmap when one is)."""
    import torch
    import oracle
    from gsx import density
    rng = np.random.default_rng(11)
    blob = rng.normal(0, 1.5, (40_000, 3))
    xyz = np.r_[blob, rng.normal(0, 1.5, (25_000, 3)) + 900.0, rng.uniform(-20, 920, (5_000, 3))].astype(np.float32)
    x = torch.from_numpy(xyz).to(cuda)
    for multi in (True, False):
        want, info_o = oracle.density_mask(xyz, voxel_size=1.0, threshold_percentage=0.05, keep_multicluster=multi)
        got, info = density.density_filter(x, 1.0, 0.05, keep_multicluster=multi)
        assert info["clusters"] == info_o["clusters"] and info["clusters"] >= 1
        assert np.array_equal(got.cpu().numpy(), want), multi
    # explicit keep lists straight into gsx_density_member_mask: a compact set (bitmap), the same set plus one far
    # voxel (hash), with points on voxel borders and outside the box on every side
    pts = np.r_[rng.uniform(-3, 12, (20_000, 3)), np.array([[5000.5, 5000.5, 5000.5], [-1e6, 3, 3], [3, 1e6, 3]])]
    pts = np.r_[pts, np.floor(rng.uniform(-3, 12, (2_000, 3)))].astype(np.float32)
    keep = np.unique(rng.integers(0, 9, (300, 3)), axis=0).astype(np.int64)
    vox = np.floor(pts / np.float32(1.0)).astype(np.int64)
    for extra in (np.zeros((0, 3), np.int64), np.array([[5000, 5000, 5000]], np.int64)):
        k = np.r_[keep, extra]
        want = (vox[:, None, :] == k[None, :, :]).all(-1).any(-1)
        got = density.member_mask(torch.from_numpy(pts).to(cuda), 1.0, k)
        assert np.array_equal(got.cpu().numpy().astype(bool), want), len(extra)


def test_bbox_alpha_match_oracle(cuda, gsx_lib):
    import torch
    import oracle
    from gsx import masks, synth
    for n in (1, 3, 1001, 200_000):
        xyz = synth.xyz(n, "mixed")
        op = synth.attributes(n)["opacity"]
        x = torch.from_numpy(xyz).to(cuda)
        for box in ((-2, -2, -2, 2, 2, 2), (-11, -11, -11, 11, 11, 11), (0.1, -0.3, 0.7, 0.1000001, 5, 9)):
            want = oracle.bbox_mask(xyz[:, 0], xyz[:, 1], xyz[:, 2], *box)
            assert np.array_equal(masks.bbox_mask(x, *box).cpu().numpy(), want)
        o = torch.from_numpy(op).to(cuda)
        for m in (1, 5, 128, 254):
            assert np.array_equal(masks.alpha_mask(o, m).cpu().numpy(), oracle.alpha_mask(op, m))


def test_density_staged_grid_equals_one_shot(cuda, gsx_lib):
    """The staged API used by the sharded driver (count into a caller grid, extract dense voxels)."""
    import torch
    from gsx import density, synth
    xyz = synth.xyz(300_000, "mixed")
    x = torch.from_numpy(xyz).to(cuda)
    voxel, thr = density.slider(0.5)
    mm = torch.cat([x.min(0).values, x.max(0).values]).cpu().numpy()
    q0, dim = density.voxel_range(mm, voxel)
    grid = torch.zeros(int(np.prod(dim)), dtype=torch.int32, device=cuda)
    half = len(xyz) // 2
    for part in (x[:half].contiguous(), x[half:].contiguous()):   # two "ranks" into one grid == all-reduce(sum)
        oob = density.grid_count(part, voxel, q0, dim, grid)
        assert int(oob.item()) == 0
    mp_ = int(len(xyz) * (thr / 100.0))
    vox, cnt, nuniq = density.grid_dense(grid, q0, dim, mp_, len(xyz))
    vox2, cnt2, nuniq2, _ = density.dense_voxels(x, voxel, mp_)
    assert np.array_equal(vox, vox2) and np.array_equal(cnt, cnt2) and nuniq == nuniq2
