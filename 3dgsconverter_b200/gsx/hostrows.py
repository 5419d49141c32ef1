"""Host-resident records: the two NumPy passes the reference spends its time in once the masks are cheap -- the
`np.column_stack` of the filter columns (data_processor.py:38,139) and the `vertices[mask]` gather of the 248-byte
records (:114,149,209,224) -- on several CPU threads inside libgsx (gsx_host_extract_xyz_opacity, gsx_host_gather_rows).
Same bytes as NumPy; arrays the C side cannot address as packed rows fall back to NumPy."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import lib, check


def _packed_rows(a: np.ndarray) -> bool:
    return isinstance(a, np.ndarray) and a.ndim == 1 and a.dtype.names is not None and a.flags.c_contiguous and \
        a.dtype.itemsize > 0 and not a.dtype.hasobject


def take_rows(a: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """a[idx] for a 1-D structured (or any fixed-itemsize) array and int64 row numbers."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    if not (isinstance(a, np.ndarray) and a.ndim == 1 and a.flags.c_contiguous and not a.dtype.hasobject
            and a.dtype.itemsize > 0) or len(idx) == 0:
        return a[idx]
    out = np.empty(len(idx), dtype=a.dtype)
    check(lib.gsx_host_gather_rows(a.ctypes.data, len(a), a.dtype.itemsize, idx.ctypes.data, len(idx), out.ctypes.data),
          "gsx_host_gather_rows")
    return out


def xyz_opacity(v: np.ndarray):
    """(np.column_stack((v['x'], v['y'], v['z'])) as float32 [n,3], v['opacity'] as float32 [n] or None)."""
    names = v.dtype.names or ()
    has_op = "opacity" in names
    f32 = np.dtype(np.float32)
    ok = _packed_rows(v) and all(nm in names and v.dtype.fields[nm][0] == f32 for nm in ("x", "y", "z")) and \
        (not has_op or v.dtype.fields["opacity"][0] == f32)
    if not ok or len(v) == 0:
        xyz = np.column_stack((v["x"], v["y"], v["z"]))
        return xyz, (v["opacity"] if has_op else None)
    n = len(v)
    xyz = np.empty((n, 3), dtype=np.float32)
    op = np.empty(n, dtype=np.float32) if has_op else None
    off = {nm: v.dtype.fields[nm][1] for nm in ("x", "y", "z")}
    check(lib.gsx_host_extract_xyz_opacity(v.ctypes.data, n, v.dtype.itemsize, off["x"], off["y"], off["z"],
                                           v.dtype.fields["opacity"][1] if has_op else -1, xyz.ctypes.data,
                                           op.ctypes.data if has_op else None), "gsx_host_extract_xyz_opacity")
    return xyz, op
