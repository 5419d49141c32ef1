"""One-off (~6 min of CPU): the reference's OWN `_kmeans_taichi` source (k_means_assign / k_means_update, gpu_ops.py:57-96,
178-191) under the serial `taichi` stand-in on a problem of the C3 SHAPE -- D = 45 (the SOG shN block), K = 256 (the
codebook size of --compression_level 5), the shape the tensor-core assign kernel is built for -- with N = 20 011 rows
(not a multiple of the 128-row MMA tile) and 2 Lloyd iterations.  Values are multiples of 1/64 (exact float32, ties
between centroids do occur).  Asserts the oracle equal bit for bit; writes tests/golden/g6_reference_kmeans_c3shape.npz
in the key layout of g4 (read by tests/test_reference_kernels_pin.py).

    python tests/golden/make_taichi_golden_c3shape.py
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


def main():
    ref = ti_serial.import_reference_gpu_ops()
    rng = np.random.default_rng(4242)
    N, D, K, iters, seed = 20_011, 45, 256, 2, 31337
    proto = rng.normal(0, 0.6, (300, D))
    X = proto[rng.integers(0, 300, N)] + rng.normal(0, 0.12, (N, D))
    X = (np.round(X * 64) / 64).astype(np.float32)
    t0 = time.time()
    np.random.seed(seed)
    C, L = ref._kmeans_taichi(X, K, max_iter=iters)
    print(f"reference kernels under the stand-in: {time.time() - t0:.0f} s", flush=True)
    np.random.seed(seed)
    rows = np.random.choice(N, K, replace=False)
    Co, Lo, cnt = oracle.kmeans_lloyd(X, K, iters, init=X[rows])
    assert np.array_equal(Lo, L), "oracle labels != reference kernels"
    assert np.array_equal(Co.view(np.uint32), C.view(np.uint32)), "oracle centroids != reference kernels"
    name = "c3shape"
    out = {f"km_{name}_X": X, f"km_{name}_meta": np.array([K, iters, seed], dtype=np.int64),
           f"km_{name}_init_rows": rows.astype(np.int64), f"km_{name}_centroids": C, f"km_{name}_labels": L}
    np.savez_compressed(HERE / "g6_reference_kmeans_c3shape.npz", **out)
    print("wrote g6_reference_kmeans_c3shape.npz; empty clusters:", int((cnt == 0).sum()))


if __name__ == "__main__":
    main()
