"""CPU: the threaded host-side row movement of libgsx (gsx_host_gather_rows / gsx_host_extract_xyz_opacity) against
NumPy -- same bytes -- on the reference's record layout (62 float32 fields, structures.py:23-59), odd layouts, thread
counts, empty inputs, bad indices; and the NumPy fallbacks of the wrappers."""
import numpy as np
import pytest


def _records(n, rng, extra=()):
    names = (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)] +
             ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])
    a = np.zeros(n, dtype=[(nm, "f4") for nm in names] + list(extra))
    for nm in a.dtype.names:
        a[nm] = rng.normal(size=n).astype(a.dtype.fields[nm][0])
    return a


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_take_rows_and_xyz_opacity_equal_numpy(threads, monkeypatch, gsx_lib):
    from gsx import hostrows
    monkeypatch.setenv("GSX_HOST_THREADS", threads)
    rng = np.random.default_rng(int(threads))
    for n in (0, 1, 7, 100_003):
        a = _records(n, rng)
        assert a.dtype.itemsize == 248
        idx = np.flatnonzero(rng.random(n) < 0.37)
        got = hostrows.take_rows(a, idx)
        assert got.dtype == a.dtype and got.tobytes() == a[idx].tobytes()
        perm = rng.permutation(n)[: n // 2]                        # any order, repeats allowed
        perm = np.r_[perm, perm[:5]]
        assert hostrows.take_rows(a, perm).tobytes() == a[perm].tobytes()
        xyz, op = hostrows.xyz_opacity(a)
        assert xyz.dtype == np.float32 and xyz.shape == (n, 3)
        assert xyz.tobytes() == np.column_stack((a["x"], a["y"], a["z"])).tobytes()
        assert op.tobytes() == np.ascontiguousarray(a["opacity"]).tobytes()


def test_odd_layouts_and_fallbacks(gsx_lib):
    from gsx import hostrows
    rng = np.random.default_rng(5)
    # unaligned float32 fields behind a 1-byte field, extra uint8 colour fields, no opacity
    a = np.zeros(5001, dtype=[("tag", "u1"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1")])
    for nm in ("x", "y", "z"):
        a[nm] = rng.normal(size=len(a)).astype(np.float32)
    a["tag"] = rng.integers(0, 255, len(a))
    assert a.dtype.itemsize == 14
    xyz, op = hostrows.xyz_opacity(a)
    assert op is None and xyz.tobytes() == np.column_stack((a["x"], a["y"], a["z"])).tobytes()
    idx = np.arange(0, len(a), 3)
    assert hostrows.take_rows(a, idx).tobytes() == a[idx].tobytes()
    # float64 coordinates and a strided view: NumPy fallbacks, same values as the reference's expressions
    b = np.zeros(100, dtype=[("x", "f8"), ("y", "f8"), ("z", "f8"), ("opacity", "f4")])
    b["x"] = rng.normal(size=100)
    xyz, op = hostrows.xyz_opacity(b)
    assert xyz.dtype == np.float64 and np.array_equal(xyz, np.column_stack((b["x"], b["y"], b["z"])))
    c = _records(200, rng)[::2]
    assert not c.flags.c_contiguous
    xyz, op = hostrows.xyz_opacity(c)
    assert xyz.tobytes() == np.column_stack((c["x"], c["y"], c["z"])).tobytes()
    assert hostrows.take_rows(c, np.array([3, 1, 1])).tobytes() == c[[3, 1, 1]].tobytes()
    # plain (non-structured) 1-D arrays work too
    d = rng.normal(size=1000)
    assert np.array_equal(hostrows.take_rows(d, np.array([5, 999, 0])), d[[5, 999, 0]])


def test_bad_indices_are_an_error_not_a_wild_read(gsx_lib):
    from gsx import hostrows
    from gsx._abi import GsxError
    a = _records(1000, np.random.default_rng(1))
    for bad in (1000, -1, 1 << 40):
        with pytest.raises(GsxError, match="outside"):
            hostrows.take_rows(a, np.array([1, 2, bad, 3], dtype=np.int64))


def test_dataprocessor_uses_the_threaded_paths_transparently(gsx_lib, monkeypatch):
    """`.data` after a filter == NumPy's vertices[idx] (the chain is stubbed: no GPU here)."""
    from gsconverter.processing import data_processor as dpm
    a = _records(20_000, np.random.default_rng(2))
    keep = np.flatnonzero(a["x"] > 0.1)

    class Chain:
        count = len(keep)
        idx = True

        def indices(self):
            return keep.astype(np.int64)

        def rebase(self):
            pass

    dp = dpm.DataProcessor(a)
    dp._chain, dp._pending = Chain(), True
    assert dp.data.tobytes() == a[keep].tobytes()
