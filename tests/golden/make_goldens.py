"""Generate tests/golden/*.npz -- run HERE (container with /root/reference), never on the GPU box.

    python tests/golden/make_goldens.py [--with-10m]

What pins what:
  * density / alpha / bbox masks and the cKDTree-SOR mean distances come from the IMPORTED
    reference itself (/root/reference, with a stub `plyfile` module -- SURVEY §8c); the script asserts
    that oracle/ reproduces them exactly before writing.
  * the Taichi-semantics SOR and the Lloyd K-Means at THESE sizes are oracle outputs (taichi is not installable;
    cross-checked against the anchor counts of SURVEY §8(c), produced by an independent numba restatement).  The
    oracle itself is pinned to the reference's kernel source on small clouds by make_taichi_goldens.py
    (serial `taichi` stand-in, g4_reference_kernels.npz).
Large arrays are stored as SHA-256 digests plus head/tail samples to keep the fixtures small.
"""
import hashlib
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))

import oracle  # noqa: E402
from gsx import synth  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def import_reference():
    sys.path.insert(0, "/root/reference")
    m = types.ModuleType("plyfile")
    m.PlyData = m.PlyElement = object
    sys.modules["plyfile"] = m
    from gsconverter.processing.data_processor import DataProcessor
    from gsconverter.processing import gpu_ops
    assert gpu_ops.HAS_TAICHI is False
    return DataProcessor, gpu_ops


def ref_filter_mask(DataProcessor, rec, call):
    """Run one reference filter on a record array tagged with the row id; return the keep-mask."""
    dp = DataProcessor(rec.copy())
    call(dp)
    mask = np.zeros(len(rec), dtype=bool)
    mask[dp.data["rid"].astype(np.int64)] = True
    return mask


def main():
    DataProcessor, ref_gpu_ops = import_reference()
    n = 100_000
    xyz = synth.xyz(n, "mixed")
    at = synth.attributes(n)
    rec = np.zeros(n, dtype=[("x", "f4"), ("y", "f4"), ("z", "f4"), ("opacity", "f4"), ("rid", "f8")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["opacity"] = at["opacity"]
    rec["rid"] = np.arange(n)
    out = {"n": n, "xyz_sha": sha(xyz), "opacity_sha": sha(at["opacity"])}

    # ---- (iii) density: the reference itself is the pin
    for sens in (0.1, 0.5, 0.9):
        for multi in (False, True):
            want = ref_filter_mask(DataProcessor, rec, lambda dp: dp.apply_density_filter(sensitivity=sens,
                                                                                            keep_multicluster=multi))
            got, _ = oracle.density_mask(xyz, sensitivity=sens, keep_multicluster=multi)
            assert np.array_equal(got, want), ("density oracle != reference", sens, multi)
            out[f"density_s{sens}_m{int(multi)}"] = np.packbits(want)
    want = ref_filter_mask(DataProcessor, rec, lambda dp: dp.apply_density_filter(0.7, 0.05, None, True))
    got, _ = oracle.density_mask(xyz, 0.7, 0.05, None, True)
    assert np.array_equal(got, want)
    out["density_v0.7_t0.05_m1"] = np.packbits(want)

    # ---- (iv) alpha / bbox: the reference itself is the pin
    for m in (1, 5, 128):
        want = ref_filter_mask(DataProcessor, rec, lambda dp: dp.apply_alpha_filter(m))
        assert np.array_equal(oracle.alpha_mask(at["opacity"], m), want)
        out[f"alpha_{m}"] = np.packbits(want)
    box = (-2, -2, -2, 2, 2, 2)
    want = ref_filter_mask(DataProcessor, rec, lambda dp: dp.crop_by_bbox(*box))
    assert np.array_equal(oracle.bbox_mask(xyz[:, 0], xyz[:, 1], xyz[:, 2], *box), want)
    out["bbox_2"] = np.packbits(want)

    # ---- (i) cKDTree semantics (data_processor.py:155-180).  The reference discards its mask (F5), so
    # the pin is its arithmetic: same SciPy calls; we check the oracle's row-mean restatement vs np.mean.
    for k, sigmas in ((27, (oracle.sor_slider(5)[1],)), (16, (1.0, 2.0, 3.0))):
        md = oracle.sor_ckdtree_mean_dists(xyz, k)
        out[f"ckd_k{k}_sha"] = sha(md)
        if k == 16:
            out["ckd_k16_means"] = md
        for s in sigmas:
            out[f"ckd_k{k}_s{s:.3f}_mask"] = np.packbits(oracle.threshold_mask(md, s))
    assert oracle.sor_slider(5) == (27, 20.0 - 4 * (17.0 / 9))
    dp = DataProcessor(rec.copy())
    dp.remove_flyers(intensity=5)  # runs the reference CPU path end to end (returns data unfiltered, F5)
    assert len(dp.data) == n

    # ---- (ii) Taichi semantics (oracle; anchors of SURVEY §8c asserted)
    anchors = {(27, "i32wrap"): {12.444: 5}, (16, "i32wrap"): {1.0: 19331, 2.0: 247, 3.0: 153}}
    for k, sigmas in ((27, (oracle.sor_slider(5)[1],)), (16, (1.0, 2.0, 3.0))):
        m32 = oracle.sor_taichi_mean_dists(xyz, k, "i32wrap")
        m64 = oracle.sor_taichi_mean_dists(xyz, k, "i64")
        assert np.array_equal(m32, m64)  # grid 15^3: no int32 overflow yet
        out[f"tai_k{k}_sha"] = sha(m32)
        if k == 16:
            out["tai_k16_means"] = m32
        for s in sigmas:
            mask = oracle.threshold_mask(m32, s)
            out[f"tai_k{k}_s{s:.3f}_mask"] = np.packbits(mask)
            exp = anchors[(k, "i32wrap")].get(round(s, 3))
            assert exp is None or int((~mask).sum()) == exp, (k, s, int((~mask).sum()))
    np.savez_compressed(HERE / "g1_100k.npz", **out)
    print("wrote g1_100k.npz")

    # ---- 1M: both hash modes diverge; digests + anchor counts
    xyz1 = synth.xyz(1_000_000, "mixed")
    g = {"xyz_sha": sha(xyz1)}
    exp = {"i32wrap": {2.0: 15932, 3.0: 15190}, "i64": {2.0: 1385, 3.0: 829}}
    for mode in ("i32wrap", "i64"):
        md = oracle.sor_taichi_mean_dists(xyz1, 16, mode)
        g[f"tai_k16_{mode}_sha"] = sha(md)
        g[f"tai_k16_{mode}_head"] = md[:1000]
        g[f"tai_k16_{mode}_tail"] = md[-1000:]
        for s in (2.0, 3.0):
            mask = oracle.threshold_mask(md, s)
            assert int((~mask).sum()) == exp[mode][s]
            g[f"tai_k16_{mode}_s{s:.1f}_mask_sha"] = sha(np.packbits(mask))
    np.savez_compressed(HERE / "g1_1m.npz", **g)
    print("wrote g1_1m.npz")

    # ---- G2: K-Means (oracle with the reference's RNG call for the init, gpu_ops.py:182)
    X45 = synth.attributes(n)["f_rest"]
    X1 = synth.attributes(50_000)["scale"].reshape(-1, 1)[:50_000].copy()
    km = {}
    for name, X, k, it in (("sh45_k16_it10", X45, 16, 10), ("sh45_k256_it10", X45, 256, 10),
                           ("scale1_k256_it20", X1, 256, 20)):
        np.random.seed(1234)
        C, L, cnt = oracle.kmeans_lloyd(X, k, it)
        np.random.seed(1234)
        init = oracle.kmeans_reference_init(X, k)
        km[f"{name}_init"] = init
        km[f"{name}_C"] = C
        km[f"{name}_counts"] = cnt
        km[f"{name}_labels_sha"] = sha(L)
        km[f"{name}_X_sha"] = sha(X)
    np.savez_compressed(HERE / "g2_kmeans.npz", **km)
    print("wrote g2_kmeans.npz")

    if "--with-10m" in sys.argv:
        xyz10 = synth.xyz(10_000_000, "mixed")
        g3 = {"xyz_sha": sha(xyz10)}
        for mode in ("i32wrap", "i64"):
            md, v = oracle.sor_taichi_mean_dists(xyz10, 16, mode, want_visits=True)
            g3[f"tai_k16_{mode}_sha"] = sha(md)
            g3[f"tai_k16_{mode}_head"] = md[:1000]
            g3[f"tai_k16_{mode}_tail"] = md[-1000:]
            g3[f"tai_k16_{mode}_visits"] = np.int64(v.sum())
            ms = np.array([np.mean(md), np.std(md)], np.float32)
            g3[f"tai_k16_{mode}_meanstd"] = ms
            for s in (2.0,):
                mask = oracle.threshold_mask(md, s)
                g3[f"tai_k16_{mode}_s{s:.1f}_mask_sha"] = sha(np.packbits(mask))
                g3[f"tai_k16_{mode}_s{s:.1f}_removed"] = np.int64((~mask).sum())
            print("10M", mode, "done", flush=True)
        dm, info = oracle.density_mask(xyz10, sensitivity=0.5, keep_multicluster=True)
        g3["density_s0.5_m1_mask_sha"] = sha(np.packbits(dm))
        g3["density_s0.5_m1_kept"] = np.int64(dm.sum())
        np.savez_compressed(HERE / "g3_10m.npz", **g3)
        print("wrote g3_10m.npz")


if __name__ == "__main__":
    main()
