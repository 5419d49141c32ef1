"""One-off (~1 h of CPU): BASELINE configs[0] -- the 100 k `mixed` cloud, `--sor_intensity 5` (k=27, sigma=12.44) and
k=16 / sigma=2 -- through the reference's OWN filter_sor_gpu source under the serial `taichi` stand-in (ti_serial.py).
Asserts the oracle equal bit for bit and writes tests/golden/g5_reference_sor_100k.npz (digests + removed counts; the
full k=16 mean-distance vector is already in g1_100k.npz and must carry the same digest).

    python tests/golden/make_taichi_golden_100k.py
"""
import hashlib
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
from make_taichi_goldens import NumpyTap  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = ti_serial.import_reference_gpu_ops()
    tap = NumpyTap()
    ref.np = tap
    xyz = synth.xyz(100_000, "mixed")
    out = {"n": 100_000, "xyz_sha": sha(xyz)}
    for k, sigma in ((27, oracle.sor_slider(5)[1]), (16, 2.0)):
        t0 = time.time()
        mask = ref.filter_sor_gpu(xyz.copy(), k=k, threshold_factor=sigma)
        means = tap.final_means
        want = oracle.sor_taichi_mean_dists(xyz, k, "i32wrap")
        assert np.array_equal(want.view(np.uint32), means.view(np.uint32)), ("oracle != reference kernel", k)
        assert np.array_equal(oracle.threshold_mask(want, sigma), mask), k
        out[f"k{k}_sigma"] = sigma
        out[f"k{k}_means_sha"] = sha(means)
        out[f"k{k}_mask_sha"] = sha(np.packbits(mask))
        out[f"k{k}_removed"] = int((~mask).sum())
        print(f"k={k} sigma={sigma:.3f}: removed {int((~mask).sum())}, {time.time() - t0:.0f} s", flush=True)
        np.savez_compressed(HERE / "g5_reference_sor_100k.npz", **out)
    print("wrote g5_reference_sor_100k.npz")


if __name__ == "__main__":
    main()
