"""One-off (~20-30 min of CPU): the reference's `sor_compute_mean_dists` KERNEL source on the 1 M `mixed` cloud -- the size
where the int32-wrapping probe hash of the kernel and the int64 hash of the host table diverge on a natural cloud
(15 932 vs 1 385 splats removed) -- for a SAMPLE of the queries: the stand-in runs ~2 us per candidate visit and the
full cloud would take ~9 h, so the reference's own host driver (`filter_sor_gpu`, unmodified: grid, int64 hash, argsort,
unique, tables over all 1 M points) is run with its kernel launch clipped to the first M = 30 000 rows of the hash-sorted
order (a pseudo-random 3 % of the cells; the kernel's `N` is only its loop bound, the table size is the separate
`hash_size` argument).  The clipped launch is the ONLY intervention.  Output: for each processed row its coordinates and
the kernel's mean distance; asserted bit-identical to the oracle's value for a point with those coordinates.

    python tests/golden/make_taichi_golden_1m_sample.py
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
from gsx import synth  # noqa: E402

M = 30_000
K = 16


def main():
    ref = ti_serial.import_reference_gpu_ops()
    real_kernel = ref.sor_compute_mean_dists
    seen = {}

    def clipped(pos, cell_start, cell_count, mean_dists, bx, by, bz, cell, hash_size, N, Kk):
        seen.update(pos=pos, mean_dists=mean_dists, hash_size=hash_size, N=N, K=Kk, cell=cell)
        real_kernel(pos, cell_start, cell_count, mean_dists, bx, by, bz, cell, hash_size, min(N, M), Kk)

    ref.sor_compute_mean_dists = clipped
    xyz = synth.xyz(1_000_000, "mixed")
    t0 = time.time()
    ref.filter_sor_gpu(xyz.copy(), k=K, threshold_factor=2.0)     # (its mask is meaningless: most rows were not processed)
    print(f"reference host driver + clipped kernel: {time.time() - t0:.0f} s", flush=True)
    assert seen["hash_size"] == 1_000_000 and seen["N"] == 1_000_000 and seen["K"] == K
    rows = np.ascontiguousarray(seen["pos"][:M])
    got = np.ascontiguousarray(seen["mean_dists"][:M])
    want32 = oracle.sor_taichi_mean_dists(xyz, K, "i32wrap")
    want64 = oracle.sor_taichi_mean_dists(xyz, K, "i64")
    # match the processed rows to original points by coordinates (points with equal coordinates have equal means)
    key = lambda a: np.ascontiguousarray(a).view([("", a.dtype)] * 3).ravel()   # noqa: E731
    order = np.argsort(key(xyz), kind="stable")
    pos_in_sorted = np.searchsorted(key(xyz)[order], key(rows))
    idx = order[pos_in_sorted]
    assert np.array_equal(xyz[idx], rows)
    assert np.array_equal(got.view(np.uint32), want32[idx].view(np.uint32)), "oracle (i32wrap) != reference kernel"
    differ = int((want32[idx].view(np.uint32) != want64[idx].view(np.uint32)).sum())
    print(f"{M} queries of the 1 M cloud: reference kernel == oracle(i32wrap) bit for bit; {differ} of them differ from the "
          f"int64-hash reading; {int((got == 0).sum())} have no candidate at all (mean 0.0)")
    np.savez_compressed(HERE / "g7_reference_sor_1m_sample.npz", rows=rows, means=got, k=K, n=1_000_000, m=M,
                        differ_from_i64=differ)
    print("wrote g7_reference_sor_1m_sample.npz")


if __name__ == "__main__":
    main()
