"""GPU: the staged pageable-host <-> HBM copies behind every *_host entry point (csrc/gsx_hostcopy.cu):
byte-exact round trips around the chunk / threshold boundaries, pinned sources on the plain path, back-to-back copies
that reuse the pinned chunks, and the SOR / K-Means host entries (pageable input == device-tensor input)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MIB = 1 << 20


@pytest.mark.parametrize("nbytes", [0, 1, 4095, 8 * MIB - 1, 8 * MIB, 8 * MIB + 1, 12 * MIB + 13, 64 * MIB,
                                    (4 * 8 * 2 + 1) * MIB + 7, 301 * MIB + 5])
def test_round_trip_bytes(nbytes, cuda, gsx_lib):
    import torch
    from gsx import hostcopy
    rng = np.random.default_rng(nbytes % 1000 + 1)
    a = rng.integers(0, 256, nbytes, dtype=np.uint8)
    t = hostcopy.to_device(a, cuda)
    assert t.shape == (nbytes,) and t.dtype == torch.uint8
    assert np.array_equal(t.cpu().numpy(), a)                    # staged upload vs torch's download
    b = hostcopy.to_host(torch.from_numpy(a).to(cuda))           # torch's upload vs staged download
    assert np.array_equal(b, a)
    assert np.array_equal(hostcopy.to_host(t), a)


def test_back_to_back_copies_reuse_chunks(cuda, gsx_lib):
    """Several copies in a row without a synchronisation in between: a pinned chunk may only be refilled after its
    previous DMA has drained."""
    from gsx import hostcopy
    rng = np.random.default_rng(5)
    arrays = [rng.normal(size=(n, 3)).astype(np.float32) for n in (3_000_000, 1_000_001, 2_500_000, 700_000)]
    tensors = [hostcopy.to_device(a, cuda) for a in arrays]
    for a, t in zip(arrays, tensors):
        assert np.array_equal(hostcopy.to_host(t), a)


def test_pinned_and_small_sources_take_the_plain_path(cuda, gsx_lib):
    import torch
    from gsx import hostcopy
    p = torch.empty((5_000_000, 3), dtype=torch.float32).pin_memory()
    p.normal_()
    t = hostcopy.to_device(p.numpy(), cuda)
    torch.cuda.synchronize()
    assert torch.equal(t.cpu(), p)
    for dt in (np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_):
        a = (np.arange(1000) % 2).astype(dt)
        assert np.array_equal(hostcopy.to_host(hostcopy.to_device(a, cuda)), a)


def test_sor_host_entry_pageable_equals_device(cuda, gsx_lib):
    """gsx_sor_filter_host on a pageable 4 M cloud (48 MB up, 4 MB + 16 MB down) == the device-tensor entry."""
    import torch
    from gsx import sor, synth
    xyz = synth.xyz(4_000_000, "mixed")
    m_h, md_h = sor.sor_filter_host(xyz, 16, 2.0, hash_mode="i32wrap", return_means=True)
    m_d, md_d = sor.sor_filter(torch.from_numpy(xyz).to(cuda), 16, 2.0, hash_mode="i32wrap", return_means=True)
    assert np.array_equal(md_h.view(np.uint32), md_d.cpu().numpy().view(np.uint32))
    assert np.array_equal(m_h, m_d.cpu().numpy())


def test_kmeans_host_entry_pageable_equals_device(cuda, gsx_lib):
    import torch
    from gsx import kmeans as gk
    rng = np.random.default_rng(9)
    X = rng.normal(0, 0.2, (400_000, 45)).astype(np.float32)     # 72 MB pageable
    init = X[rng.choice(len(X), 256, replace=False)]
    C_h, L_h = gk.kmeans_host(X, 256, 2, init)
    C_d, L_d, _ = gk.kmeans_lloyd(torch.from_numpy(X).to(cuda), 256, 2, init=torch.from_numpy(init).to(cuda))
    assert np.array_equal(L_h, L_d.cpu().numpy())
    assert np.array_equal(C_h.view(np.uint32), C_d.cpu().numpy().view(np.uint32))
