"""Generate tests/golden/g4_reference_kernels.npz -- run HERE (container with /root/reference), never on the GPU box.

    python tests/golden/make_taichi_goldens.py

Executes the reference's OWN, UNMODIFIED `gsconverter/processing/gpu_ops.py` -- `filter_sor_gpu` (host grid build +
the `sor_compute_mean_dists` kernel + mean/std threshold) and `_kmeans_taichi` (`np.random.choice` init + the
`k_means_assign` / `k_means_update` kernels) -- with tests/golden/ti_serial.py standing in for the `taichi` module
(the real wheel is not installable here; the stand-in's five stated assumptions T1-T5 are in its header).  The
outputs are therefore produced by the reference's source text, not by a restatement; this script asserts that
oracle/ reproduces every one of them bit for bit before writing.  Small clouds only: the stand-in runs the kernel
bodies as Python (~2.5 ms per point).
"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))

import oracle  # noqa: E402
import ti_serial  # noqa: E402


class NumpyTap:
    """`np` as seen by the reference module: records the vector handed to np.std (= final_means, gpu_ops.py:259)."""

    def __init__(self):
        self.final_means = None

    def __getattr__(self, name):
        return getattr(np, name)

    def std(self, a, *args, **kw):
        self.final_means = np.array(a, copy=True)
        return np.std(a, *args, **kw)


def sor_clouds():
    """name -> (xyz float32 [N,3], list of (k, sigma)).  The generator below is the committed definition."""
    rng = np.random.default_rng(20260923)
    out = {}
    out["uniform3k"] = (rng.uniform(-1, 1, (3000, 3)).astype(np.float32), [(16, 2.0)])
    # clustered + background + far flyers: long and short buckets, real outliers
    c = np.concatenate([rng.normal(0, 0.02, (1200, 3)) + [0.3, 0.2, -0.1], rng.normal(0, 0.05, (900, 3)) - [0.4, 0.1, 0.2],
                        rng.uniform(-1, 1, (860, 3)), rng.uniform(-6, 6, (40, 3))])
    out["mixed3k"] = (c.astype(np.float32), [(1, 1.0), (16, 2.0), (27, 12.444), (80, 3.0)])   # 80 -> capped at 50
    # 2000 x 1 x 1 slab: grid indices up to ~900, so nx * 73856093 wraps int32 (T1) and probes alias other buckets
    e = rng.uniform(0, 1, (3000, 3)) * [2000.0, 1.0, 1.0]
    out["elongated_i32wrap"] = (e.astype(np.float32), [(16, 2.0)])
    # the same uniform cube far from the origin: float32 cancellation in (p - bbox_min) / cell_size
    out["offset1e4"] = ((rng.uniform(-1, 1, (2000, 3)) + [1.0e4, -2.0e4, 3.0e4]).astype(np.float32), [(16, 1.0)])
    # exact duplicates (d2 == 0 is skipped as "self", gpu_ops.py:152) incl. triples
    base = rng.uniform(-1, 1, (900, 3)).astype(np.float32)
    out["duplicates"] = (np.concatenate([base, base[:600], base[:150]]), [(16, 2.0)])
    # planar cloud: extent z == 0 -> vol <= 0 -> vol = 1.0 (gpu_ops.py:204)
    p = rng.uniform(-1, 1, (1500, 3))
    p[:, 2] = 0.25
    out["planar"] = (p.astype(np.float32), [(16, 2.0)])
    out["tiny5"] = (rng.uniform(-1, 1, (5, 3)).astype(np.float32), [(16, 1.0)])          # k > N: sentinels stay
    out["identical64"] = (np.full((64, 3), 0.5, dtype=np.float32), [(16, 1.0)])          # nothing but "self" hits
    return out


def kmeans_cases():
    """name -> (X float32 [N,D], K, iterations, numpy legacy seed)."""
    rng = np.random.default_rng(77)
    out = {}
    out["sh45"] = ((rng.normal(0, 0.1, (2000, 45)) + rng.integers(0, 5, (2000, 1)) * 0.2).astype(np.float32), 16, 3, 1234)
    x = rng.uniform(-1, 1, (1500, 3)).astype(np.float32)
    x[100:400] = x[:300]            # duplicate rows: if two are drawn as init, the later centroid stays empty (-> 0)
    x[400:420] = x[0]
    x += np.float32(5.0)            # ... and, being far from the origin, stays empty in every later iteration
    out["dups3"] = (x, 24, 5, 0)    # seed 0 draws two identical rows
    out["codebook1d"] = (rng.normal(0, 1, (600, 1)).astype(np.float32), 64, 4, 99)
    # integer lattice: many exact distance ties -> lowest index wins (strict `<`, gpu_ops.py:71)
    out["lattice"] = (rng.integers(0, 3, (800, 4)).astype(np.float32), 12, 4, 5)
    return out


def main():
    ref = ti_serial.import_reference_gpu_ops()
    tap = NumpyTap()
    ref.np = tap
    out = {}
    t0 = time.time()
    for name, (xyz, runs) in sor_clouds().items():
        out[f"sor_{name}_xyz"] = xyz
        for k, sigma in runs:
            mask = ref.filter_sor_gpu(xyz.copy(), k=k, threshold_factor=sigma)
            means = tap.final_means
            assert means is not None and means.dtype == np.float32 and mask.dtype == bool
            want = oracle.sor_taichi_mean_dists(xyz, k, "i32wrap")
            assert np.array_equal(want.view(np.uint32), means.view(np.uint32)), ("oracle != reference kernel", name, k)
            assert np.array_equal(oracle.threshold_mask(want, sigma), mask), ("oracle mask != reference", name, k)
            out[f"sor_{name}_k{k}_s{sigma}_means"] = means
            out[f"sor_{name}_k{k}_s{sigma}_mask"] = mask
            print(f"SOR {name} k={k} sigma={sigma}: removed {int((~mask).sum())}/{len(xyz)}  [{time.time() - t0:.0f}s]",
                  flush=True)
    ref.np = np
    for name, (X, K, iters, seed) in kmeans_cases().items():
        np.random.seed(seed)
        C, L = ref._kmeans_taichi(X, K, max_iter=iters)
        np.random.seed(seed)
        init_rows = np.random.choice(len(X), K, replace=False)
        Co, Lo, cnt = oracle.kmeans_lloyd(X, K, iters, init=X[init_rows])
        assert np.array_equal(Lo, L), ("oracle labels != reference kernels", name)
        assert np.array_equal(Co.view(np.uint32), C.view(np.uint32)), ("oracle centroids != reference kernels", name)
        out[f"km_{name}_X"] = X
        out[f"km_{name}_meta"] = np.array([K, iters, seed], dtype=np.int64)
        out[f"km_{name}_init_rows"] = init_rows.astype(np.int64)
        out[f"km_{name}_centroids"] = C
        out[f"km_{name}_labels"] = L
        print(f"K-Means {name}: N={len(X)} D={X.shape[1]} K={K} it={iters}, empty clusters {int((cnt == 0).sum())}  "
              f"[{time.time() - t0:.0f}s]", flush=True)
    np.savez_compressed(HERE / "g4_reference_kernels.npz", **out)
    print("wrote", HERE / "g4_reference_kernels.npz")


if __name__ == "__main__":
    main()
