"""Synthetic splat clouds pinned by SURVEY.md §8(d) (normative draw order).

Block-wise generation so that any shard can regenerate its own part: block b
covers points [b*2**20, (b+1)*2**20) and uses ``default_rng([SEED, b])``.
Values are drawn in float64 and cast to float32.  Used by tests, goldens and
bench.py (inputs only -- nothing here is on the measured path).
"""
from __future__ import annotations

import math
import numpy as np

SEED = 20260923
BLOCK = 1 << 20


def _cluster_params():
    r0 = np.random.default_rng([SEED, 2**31 - 1])
    c = r0.uniform(-8, 8, (16, 3))
    s = 0.1 * 4 ** r0.uniform(0, 1, 16)
    return c, s


def xyz(n: int, kind: str = "mixed", start_block: int = 0) -> np.ndarray:
    """float32[n,3] cloud.  kind in {mixed, uniform, clustered}.

    ``start_block`` lets a shard generate blocks [start_block, ...) only; n is
    then the number of points of that shard (must start on a block boundary).
    """
    if kind not in ("mixed", "uniform", "clustered"):
        raise ValueError(kind)
    c, s = _cluster_params()
    out = np.empty((n, 3), dtype=np.float32)
    nb = math.ceil(n / BLOCK)
    for b in range(nb):
        m = min(BLOCK, n - b * BLOCK)
        r = np.random.default_rng([SEED, start_block + b])
        u = r.uniform(0, 1, m)
        uni = r.uniform(-10, 10, (m, 3))
        j = r.integers(0, 16, m)
        g = c[j] + s[j, None] * r.standard_normal((m, 3))
        f = r.uniform(-12, 12, (m, 3))
        if kind == "mixed":
            blk = np.where((u < 0.4975)[:, None], uni, np.where((u < 0.995)[:, None], g, f))
        elif kind == "uniform":
            blk = uni
        else:
            blk = g
        out[b * BLOCK : b * BLOCK + m] = blk.astype(np.float32)
    return out


def attributes(n: int, sh_coeffs: int = 45) -> dict:
    """Attribute columns of SURVEY §8(d): opacity (logit), scales, rots, f_dc, f_rest."""
    P = np.random.default_rng([SEED, 2**31 - 2]).normal(0, 0.15, (1024, 45))
    out = {
        "opacity": np.empty(n, np.float32),
        "scale": np.empty((n, 3), np.float32),
        "rot": np.empty((n, 4), np.float32),
        "f_dc": np.empty((n, 3), np.float32),
        "f_rest": np.empty((n, sh_coeffs), np.float32),
    }
    nb = math.ceil(n / BLOCK)
    for b in range(nb):
        m = min(BLOCK, n - b * BLOCK)
        sl = slice(b * BLOCK, b * BLOCK + m)
        r = np.random.default_rng([SEED, 2**30 + b])
        out["opacity"][sl] = r.normal(0, 2, m)
        out["scale"][sl] = r.normal(-4.5, 1, (m, 3))
        q = r.standard_normal((m, 4))
        out["rot"][sl] = q / np.linalg.norm(q, axis=1, keepdims=True)
        out["f_dc"][sl] = r.normal(0, 1, (m, 3))
        mi = r.integers(0, 1024, m)
        out["f_rest"][sl] = (P[mi] + r.normal(0, 0.03, (m, 45)))[:, :sh_coeffs]
    return out


def structured(n: int, kind: str = "mixed", sh_degree: int = 3) -> np.ndarray:
    """The reference's interchange record (structures.py:23-59 layout) filled with synthetic data."""
    nc = 3 * ((sh_degree + 1) ** 2 - 1)
    dt = [(k, "f4") for k in ("x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2")]
    dt += [(f"f_rest_{i}", "f4") for i in range(nc)]
    dt += [(k, "f4") for k in ("opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3")]
    a = np.zeros(n, dtype=dt)
    p = xyz(n, kind)
    at = attributes(n, nc)
    a["x"], a["y"], a["z"] = p[:, 0], p[:, 1], p[:, 2]
    for i in range(3):
        a[f"f_dc_{i}"] = at["f_dc"][:, i]
        a[f"scale_{i}"] = at["scale"][:, i]
    for i in range(4):
        a[f"rot_{i}"] = at["rot"][:, i]
    for i in range(nc):
        a[f"f_rest_{i}"] = at["f_rest"][:, i]
    a["opacity"] = at["opacity"]
    return a
