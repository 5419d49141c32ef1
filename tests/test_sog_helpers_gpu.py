"""GPU: SOG writer helpers against NumPy: np.lexsort((z,y,x)) and the reference's
quantize_to_codebook (formats/sog.py:264, :408-419) -- index-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref_quantize(vals, cb):
    """formats/sog.py:408-419 restated with the same NumPy calls."""
    if len(cb) == 1:
        return np.zeros_like(vals, dtype=np.uint8)
    idx = np.clip(np.searchsorted(cb, vals), 0, len(cb) - 1)
    left = np.maximum(idx - 1, 0)
    use_left = np.abs(vals - cb[left]) < np.abs(vals - cb[idx])
    idx[use_left] = left[use_left]
    return idx.astype(np.uint8)


@pytest.mark.parametrize("n", [1, 33, 4097, 1_000_003])
def test_lexsort_matches_numpy(n, cuda, gsx_lib):
    import torch
    from gsx import sog
    rng = np.random.default_rng(n)
    xyz = rng.standard_normal((n, 3)).astype(np.float32)
    if n > 100:  # ties on x and (x,y), negative zero, repeated rows: stability and key mapping
        xyz[: n // 2, 0] = np.round(xyz[: n // 2, 0] * 2) / 2
        xyz[: n // 4, 1] = np.round(xyz[: n // 4, 1])
        xyz[5:40] = xyz[45:80]
        xyz[100:110, 0] = -0.0
        xyz[110:120, 0] = 0.0
    want = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    got = sog.lexsort_zyx(torch.from_numpy(xyz).to(cuda)).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("m", [1, 2, 16, 256, 1000])
def test_quantize_to_codebook_matches_reference(m, cuda, gsx_lib):
    import torch
    from gsx import sog
    rng = np.random.default_rng(m)
    cb = np.sort(rng.normal(-4.5, 1.0, m).astype(np.float32))
    vals = rng.normal(-4.5, 1.3, 500_003).astype(np.float32)
    vals[:m] = cb                                        # exact hits
    if m > 2:
        vals[m:2 * m - 1] = (cb[:-1] + cb[1:]) / 2       # midpoints: the strict '<' tie rule
    vals[-3:] = [cb[0] - 10, cb[-1] + 10, cb[m // 2]]
    got = sog.quantize_to_codebook(torch.from_numpy(vals).to(cuda), cb).cpu().numpy()
    assert np.array_equal(got, _ref_quantize(vals, cb))


def test_codebook_1d_matches_oracle(cuda, gsx_lib):
    import oracle
    from gsx import sog, synth
    s = synth.attributes(100_000)["scale"].reshape(-1)   # 300 000 scalars > 50 000: the subsample path
    np.random.seed(5)
    cb = sog.codebook_1d(s, 256, 20)
    np.random.seed(5)
    fit = s[np.random.choice(len(s), 50000, replace=False)]
    C, _, _ = oracle.kmeans_lloyd(fit.reshape(-1, 1), 256, 20)   # consumes the kmeans init draw the same way
    assert np.array_equal(cb, np.array(sorted(C.flatten()), dtype=np.float32))
