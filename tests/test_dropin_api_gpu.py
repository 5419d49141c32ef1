"""GPU: the reference-facing plugin surface (gsconverter.processing) -- same names, argument meaning,
messages and error behaviour as the reference, results checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _records(n):
    from gsx import synth
    return synth.structured(n, "mixed")


def test_has_taichi_flag_and_module_surface(cuda, gsx_lib):
    from gsconverter.processing import DataProcessor, gpu_ops
    assert gpu_ops.HAS_TAICHI is True
    for name in ("kmeans", "filter_sor_gpu", "HAS_TAICHI"):
        assert hasattr(gpu_ops, name)
    for name in ("crop_by_bbox", "apply_alpha_filter", "apply_density_filter", "remove_flyers", "apply_auto_bbox",
                 "cap_sh_degree", "add_rgb_from_sh"):
        assert hasattr(DataProcessor, name)


def test_filter_chain_matches_oracle(cuda, gsx_lib, capsys):
    """converter.py:194-236 order: bbox -> alpha -> density -> SOR."""
    import oracle
    from gsconverter.processing import DataProcessor
    rec = _records(200_000)
    dp = DataProcessor(rec.copy())
    keep = np.ones(len(rec), bool)
    # bbox
    out = dp.crop_by_bbox(-11, -11, -11, 11, 11, 11)
    keep &= oracle.bbox_mask(rec["x"], rec["y"], rec["z"], -11, -11, -11, 11, 11, 11)
    assert out is dp.data and np.array_equal(dp.data, rec[keep])
    # alpha
    cur = rec[keep]
    dp.apply_alpha_filter(5)
    m = oracle.alpha_mask(cur["opacity"], 5)
    cur = cur[m]
    assert np.array_equal(dp.data, cur)
    # density
    dp.apply_density_filter(sensitivity=0.5, keep_multicluster=True)
    m, _ = oracle.density_mask(np.column_stack((cur["x"], cur["y"], cur["z"])), sensitivity=0.5, keep_multicluster=True)
    cur = cur[m]
    assert np.array_equal(dp.data, cur)
    # SOR via the slider (k=27, sigma=12.44) and via explicit k/sigma
    dp.remove_flyers(intensity=5)
    xyz = np.column_stack((cur["x"], cur["y"], cur["z"]))
    cur = cur[oracle.sor_taichi_mask(xyz, 27, 20.0 - 4 * (17.0 / 9))]
    assert np.array_equal(dp.data, cur)
    dp.remove_flyers(16, 2.0)
    xyz = np.column_stack((cur["x"], cur["y"], cur["z"]))
    cur = cur[oracle.sor_taichi_mask(xyz, 16, 2.0)]
    assert np.array_equal(dp.data, cur)
    text = capsys.readouterr().out
    assert "After cropping, retained" in text and "Alpha Filter (min 5): Retained" in text
    assert "After density filter, retained" in text and "After removing flyers (GPU), retained" in text


def test_edge_behaviour(cuda, gsx_lib):
    from gsconverter.processing import DataProcessor, gpu_ops
    rec = _records(5_000)
    dp = DataProcessor(rec.copy())
    assert dp.apply_alpha_filter(0) is None and len(dp.data) == 5_000          # <=0: no-op
    dp.apply_alpha_filter(255)
    assert len(dp.data) == 0                                                   # >=255: empties
    dp = DataProcessor(rec[["x", "y", "z"]].copy())
    assert dp.apply_alpha_filter(5) is None and len(dp.data) == 5_000          # no opacity field: skipped
    with pytest.raises(TypeError):
        DataProcessor([1, 2, 3]).remove_flyers()
    with pytest.raises(TypeError):
        DataProcessor([1, 2, 3]).apply_density_filter()
    with pytest.raises(ValueError):
        gpu_ops.filter_sor_gpu(np.zeros((10, 2), np.float32))
    dp = DataProcessor(rec.copy())
    dp.apply_density_filter(voxel_size=0.01, threshold_percentage=50.0)       # nothing dense -> empty
    assert len(dp.data) == 0 and dp.data.dtype == rec.dtype


def test_kmeans_plugin_semantics(cuda, gsx_lib):
    import oracle
    from gsconverter.processing import gpu_ops
    from gsx import synth
    X = synth.attributes(30_000)["f_rest"]
    np.random.seed(77)
    C, L = gpu_ops.kmeans(X, 64, max_iter=6)
    np.random.seed(77)
    Co, Lo, _ = oracle.kmeans_lloyd(X, 64, 6)          # same single np.random.choice draw (gpu_ops.py:182)
    assert C.dtype == np.float32 and L.dtype == np.int32
    assert np.array_equal(L, Lo) and np.allclose(C, Co, rtol=1e-5, atol=0)
    small = X[:10]
    C, L = gpu_ops.kmeans(small, 16)                    # k >= N: passthrough (gpu_ops.py:30-31)
    assert np.array_equal(C, small) and np.array_equal(L, np.arange(10, dtype=np.int32))
    C1, L1 = gpu_ops.kmeans(synth.attributes(50_000)["scale"].reshape(-1, 1)[:50_000], 256, max_iter=20)
    assert C1.shape == (256, 1) and L1.shape == (50_000,)


def _oracle_chain(rec):
    """bbox -> alpha -> density -> SOR (converter.py:194-236 order) with the oracle; returns surviving rows."""
    import oracle
    idx = np.arange(len(rec))
    cur = rec
    m = oracle.bbox_mask(cur["x"], cur["y"], cur["z"], -11, -11, -11, 11, 11, 11)
    idx, cur = idx[m], cur[m]
    m = oracle.alpha_mask(cur["opacity"], 5)
    idx, cur = idx[m], cur[m]
    m, _ = oracle.density_mask(np.column_stack((cur["x"], cur["y"], cur["z"])), sensitivity=0.5, keep_multicluster=True)
    idx, cur = idx[m], cur[m]
    m = oracle.sor_taichi_mask(np.column_stack((cur["x"], cur["y"], cur["z"])), 16, 2.0)
    return idx[m]


def test_deferred_compaction_gathers_once(cuda, gsx_lib):
    """defer_compaction=True (what gsx.dropin.patch sets for converter.py): filters return None, the
    248-byte records are gathered once when `.data` is read -- same rows as the eager chain."""
    from gsconverter.processing import DataProcessor
    rec = _records(300_000)
    want = _oracle_chain(rec)
    old = DataProcessor.defer_compaction
    DataProcessor.defer_compaction = True
    try:
        dp = DataProcessor(rec)
        assert dp.crop_by_bbox(-11, -11, -11, 11, 11, 11) is None
        assert dp.apply_alpha_filter(5) is None
        assert dp.apply_density_filter(1.0, 0.32, sensitivity=0.5, keep_multicluster=True) is None
        assert dp._data is rec                      # nothing gathered on the host yet
        assert dp.remove_flyers(16, 2.0) is None
        out = dp.data
        assert np.array_equal(out, rec[want])
        assert dp.data is out                       # idempotent
        dp.apply_auto_bbox()
        dp.add_rgb_from_sh()
        assert "red" in dp.data.dtype.names and len(dp.data) == len(want)
        dp.remove_flyers(16, 2.0)                   # chain still valid after add_rgb (same rows)
        assert len(dp.data) <= len(want)
    finally:
        DataProcessor.defer_compaction = old


def test_filter_chain_device_indices(cuda, gsx_lib):
    from gsx.pipeline import FilterChain
    rec = _records(300_000)
    want = _oracle_chain(rec)
    ch = FilterChain(np.column_stack((rec["x"], rec["y"], rec["z"])), rec["opacity"])
    assert np.array_equal(ch.indices(), np.arange(len(rec)))
    ch.crop_by_bbox(-11, -11, -11, 11, 11, 11)
    ch.alpha(5)
    ch.density(sensitivity=0.5, keep_multicluster=True)
    n = ch.sor(16, 2.0, hash_mode="i32wrap")
    assert n == len(want) and np.array_equal(ch.indices(), want)
    ch.crop_by_bbox(100, 100, 100, 101, 101, 101)          # removes everything
    assert ch.count == 0 and len(ch.indices()) == 0
    assert ch.sor(16, 2.0) == 0 and ch.density(sensitivity=0.5)[0] == 0


def test_compact_points_matches_numpy(cuda, gsx_lib):
    import torch
    from gsx.pipeline import compact
    rng = np.random.default_rng(4)
    for n in (1, 31, 1024, 1025, 100_003, 3_000_000):
        xyz = rng.standard_normal((n, 3)).astype(np.float32)
        op = rng.standard_normal(n).astype(np.float32)
        for p in (0.0, 0.5, 1.0):
            mask = rng.random(n) < p
            x, o, idx, m = compact(torch.from_numpy(mask).to(cuda), torch.from_numpy(xyz).to(cuda),
                                   torch.from_numpy(op).to(cuda), None)
            assert m == int(mask.sum())
            assert np.array_equal(x.cpu().numpy(), xyz[mask]) and np.array_equal(o.cpu().numpy(), op[mask])
            assert np.array_equal(idx.cpu().numpy(), np.flatnonzero(mask))


def test_compact_points_unaligned_mask_and_chained_index(cuda, gsx_lib):
    """mask at an odd byte offset (the 8-byte mask loads must fall back), an input row index (second filter of a chain)
    and no opacity column."""
    import torch
    from gsx.pipeline import compact
    rng = np.random.default_rng(11)
    n = 300_007
    xyz = rng.standard_normal((n, 3)).astype(np.float32)
    rows = rng.permutation(10 * n)[:n].astype(np.int32)
    mask = rng.random(n) < 0.3
    buf = torch.zeros(n + 3, dtype=torch.uint8, device=cuda)
    buf[3:] = torch.from_numpy(mask.view(np.uint8)).to(cuda)
    x, o, idx, m = compact(buf[3:], torch.from_numpy(xyz).to(cuda), None, torch.from_numpy(rows).to(cuda))
    assert o is None and m == int(mask.sum())
    assert np.array_equal(x.cpu().numpy(), xyz[mask]) and np.array_equal(idx.cpu().numpy(), rows[mask])
