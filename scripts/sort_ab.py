"""A/B of the radix sort forms (run once per libgsx variant, GSX_LIB=...): gsx_sort_pairs on (uint64, int32) pairs, the
keys-only form on packed words, and the whole single-GPU grid build, at 10 M and 80 M.  CUDA events, min of 5."""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))
import torch
from gsx import sor, synth


def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(min(ts), 3)


def main():
    dev = torch.device("cuda:0")
    out = {"lib": sys.argv[1] if len(sys.argv) > 1 else "default"}
    for n in (10_000_000, 80_000_000):
        g = torch.Generator(device=dev).manual_seed(1)
        bits = 39 if n <= 16_777_216 else 42
        keys0 = torch.randint(0, 1 << bits, (n,), device=dev, dtype=torch.int64, generator=g)
        vals0 = torch.arange(n, device=dev, dtype=torch.int32)
        keys, vals = keys0.clone(), vals0.clone()

        def pairs():
            keys.copy_(keys0); vals.copy_(vals0)
            sor.sort_pairs(keys, vals, 0, bits)

        def copy_only():
            keys.copy_(keys0); vals.copy_(vals0)

        idx_bits = 24 if n <= 16_777_216 else 27
        kb = min(bits, 64 - idx_bits)
        words0 = ((keys0 & ((1 << kb) - 1)) << idx_bits) | vals0.to(torch.int64)
        words = words0.clone()

        def keys_only():
            words.copy_(words0)
            sor.sort_pairs(words, None, idx_bits, idx_bits + kb)

        def copy_words():
            words.copy_(words0)

        t_pairs, t_copy = ev(pairs), ev(copy_only)
        t_keys, t_cw = ev(keys_only), ev(copy_words)
        # checks: sorted; the keys-only form agrees with a stable pair sort of the same keys
        pairs(); torch.cuda.synchronize()
        assert bool((keys[1:] >= keys[:-1]).all())
        keys_only(); torch.cuda.synchronize()
        k2 = (words0 >> idx_bits) & ((1 << kb) - 1)
        v2 = vals0.clone()
        sor.sort_pairs(k2, v2, 0, kb); torch.cuda.synchronize()
        assert bool((((words >> idx_bits) & ((1 << kb) - 1)) == k2).all())
        assert bool(((words & ((1 << idx_bits) - 1)).to(torch.int32) == v2).all())
        del keys, vals, keys0, vals0, words, words0, k2, v2
        xyz = torch.from_numpy(synth.xyz(n, "mixed")).to(dev)
        ws = sor.workspace(n, dev)
        t_build = ev(lambda: sor.build_grid(xyz, ws))
        out[str(n)] = {"key_bits": bits, "sort_pairs_ms": round(t_pairs - t_copy, 3), "keys_only_bits": kb,
                       "sort_keys_only_ms": round(t_keys - t_cw, 3), "grid_build_ms": t_build}
        del xyz, ws
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
