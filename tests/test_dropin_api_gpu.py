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
