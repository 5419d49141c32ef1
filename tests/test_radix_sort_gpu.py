"""GPU: the hand-written stable LSD radix sort (csrc/gsx_radix.cu) against torch's stable sort."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 511, 512, 513, 4095, 4096, 4097, 100_003, 1_000_000, 5_000_001])
@pytest.mark.parametrize("bits", [(0, 64), (0, 42), (0, 8), (0, 19), (8, 40)])
def test_sort_pairs_matches_stable_sort(n, bits, cuda, gsx_lib):
    import torch
    from gsx import sor
    g = torch.Generator(device=cuda).manual_seed(n * 131 + bits[1])
    b0, b1 = bits
    # few distinct values in the sorted field -> many ties -> stability is really exercised
    span = min(b1 - b0, 62)
    field = torch.randint(0, min(1 << span, 50_000 if n > 1000 else 1 << span), (n,), device=cuda, generator=g,
                          dtype=torch.int64)
    noise_lo = torch.randint(0, 1 << b0, (n,), device=cuda, generator=g, dtype=torch.int64) if b0 > 0 else 0
    keys = (field << b0) | noise_lo            # bits below begin_bit must be ignored by the sort
    vals = torch.arange(n, device=cuda, dtype=torch.int32)
    want_order = torch.sort(field, stable=True).indices
    k, v = sor.sort_pairs(keys.clone(), vals.clone(), b0, b1)
    assert torch.equal(v.long(), want_order)
    assert torch.equal(k, keys[want_order])


@pytest.mark.parametrize("n", [1, 33, 4096, 4097, 100_003, 3_000_001])
@pytest.mark.parametrize("key_bits", [8, 19, 39])
def test_sort_keys_only_packed_words(n, key_bits, cuda, gsx_lib):
    """vals=None: bare 64-bit words, payload (the element index) packed below begin_bit -- the form the grid build uses.
    Sorting only the bits above the index must give the stable order of the key field."""
    import torch
    from gsx import sor
    g = torch.Generator(device=cuda).manual_seed(n * 7 + key_bits)
    idx_bits = max(1, int(np.ceil(np.log2(max(n, 2)))))
    field = torch.randint(0, min(1 << key_bits, 40_000), (n,), device=cuda, generator=g, dtype=torch.int64)
    words = (field << idx_bits) | torch.arange(n, device=cuda, dtype=torch.int64)
    want_order = torch.sort(field, stable=True).indices
    out, _ = sor.sort_pairs(words.clone(), None, idx_bits, idx_bits + key_bits)
    assert torch.equal(out & ((1 << idx_bits) - 1), want_order)
    assert torch.equal(out >> idx_bits, field[want_order])


def test_sort_many_tiles_lookback_chain(cuda, gsx_lib):
    """All keys equal: every tile's whole count lands on one digit, so every look-back walks the full chain of
    predecessors (the longest dependency the decoupled look-back can see)."""
    import torch
    from gsx import sor
    n = 6_000_000
    keys = torch.full((n,), 0x5A5A5A5A5A, device=cuda, dtype=torch.int64)
    vals = torch.arange(n, device=cuda, dtype=torch.int32)
    k, v = sor.sort_pairs(keys.clone(), vals.clone(), 0, 40)
    assert torch.equal(v, vals) and torch.equal(k, keys)
