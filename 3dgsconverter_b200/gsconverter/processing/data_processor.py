"""Drop-in for ``gsconverter/processing/data_processor.py``: the filter API on a NumPy structured
array, with every keep-mask computed by libgsx.so on the GPU.

Public surface and behaviour follow the reference (data_processor.py:7-354): each filter rebinds
``self.data`` to the compacted array and returns it; slider formulas (:17-28, :125-134), early-outs
and console messages are preserved.  Host-only helpers (RGB from SH, SH capping, auto-bbox report)
are plain NumPy and stay on the host (SURVEY §2.1 "OUT").

No CPU fallback: if the CUDA backend is unavailable the filters raise instead of silently running
the reference's SciPy path (whose mask the reference computes and then discards -- SURVEY F5).
"""
from __future__ import annotations

import numpy as np

from ..utils.utility_functions import debug_print, status_print

_SH_C0 = 0.28209479177387814


def _xyz(vertices) -> np.ndarray:
    return np.column_stack((vertices["x"], vertices["y"], vertices["z"]))


def _to_device(a: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda")


def _require_backend():
    from .gpu_ops import HAS_TAICHI
    if not HAS_TAICHI:
        raise RuntimeError("gsx: no CUDA device available and CPU fallback is disabled by design")


class DataProcessor:
    def __init__(self, data):
        self.data = data

    # ------------------------------------------------------------------ density (:11-117)
    def apply_density_filter(self, voxel_size=1.0, threshold_percentage=0.32, sensitivity=None,
                             keep_multicluster=False):
        debug_print("[DEBUG] Executing 'apply_density_filter' function...")
        if not isinstance(self.data, np.ndarray):
            raise TypeError("self.data must be a numpy structured array.")
        _require_backend()
        from gsx import density
        if sensitivity is not None:
            voxel_size, threshold_percentage = density.slider(sensitivity)
        debug_print(f"Density Filter Params: Voxel={voxel_size:.4f}, Thresh={threshold_percentage:.4f}%, "
                    f"MultiCluster={keep_multicluster}")
        vertices = self.data
        if len(vertices) == 0:
            status_print("Warning: Density filter removed all points.")
            return self.data
        mask, info = density.density_filter(_to_device(_xyz(vertices)), voxel_size, threshold_percentage, None,
                                            keep_multicluster)
        debug_print(f"[DEBUG] Found {info['voxels']} unique voxels.")
        if info["dense"] == 0 or info["clusters"] == 0:
            status_print("Warning: Density filter removed all points.")
            self.data = self.data[:0]
            return self.data
        self.data = vertices[mask.cpu().numpy()]
        status_print(f"Density Filter: Kept {info['clusters']} clusters (largest: {info['max_len']} voxels).")
        status_print(f"After density filter, retained {len(self.data)} out of {len(vertices)} vertices.")
        return self.data

    # ------------------------------------------------------------------ SOR (:119-182)
    def remove_flyers(self, k=25, threshold_factor=10.5, chunk_size=50000, intensity=None):
        debug_print("[DEBUG] Executing 'remove_flyers' function...")
        if not isinstance(self.data, np.ndarray):
            raise TypeError("self.data must be a numpy structured array.")
        if intensity is not None:  # :125-134
            k = int(10 + (intensity - 1) * (40 / 9))
            threshold_factor = 20.0 - (intensity - 1) * (17.0 / 9)
        debug_print(f"SOR Filter (Remove Flyers) Params: K={k}, Sigma={threshold_factor:.2f}")
        vertices = self.data
        num_points = len(vertices)
        from .gpu_ops import filter_sor_gpu
        _require_backend()
        status_print("[SOR] Determining outliers on GPU (gsx / sm_100a)...")
        if num_points == 0:
            return self.data
        gpu_mask = filter_sor_gpu(_xyz(vertices), k, threshold_factor, verbose=True)
        if gpu_mask is None:
            raise RuntimeError("gsx: SOR backend returned no mask")
        self.data = vertices[gpu_mask]
        status_print(f"After removing flyers (GPU), retained {len(self.data)} out of {num_points} vertices.")
        return self.data

    # ------------------------------------------------------------------ alpha (:184-213)
    def apply_alpha_filter(self, min_opacity_u8):
        debug_print(f"[DEBUG] Executing 'apply_alpha_filter' with min={min_opacity_u8}")
        if "opacity" not in self.data.dtype.names:
            status_print("Warning: No opacity channel found. Alpha filter skipped.")
            return
        limit = min_opacity_u8
        if limit <= 0:
            return
        if limit >= 255:
            self.data = self.data[:0]
            return
        _require_backend()
        from gsx import masks
        original_len = len(self.data)
        if original_len:
            keep = masks.alpha_mask(_to_device(self.data["opacity"]), limit).cpu().numpy()
            self.data = self.data[keep]
        status_print(f"Alpha Filter (min {limit}): Retained {len(self.data)} out of {original_len} splats.")
        return self.data

    # ------------------------------------------------------------------ bbox (:215-231)
    def crop_by_bbox(self, min_x, min_y, min_z, max_x, max_y, max_z):
        _require_backend()
        from gsx import masks
        if len(self.data):
            keep = masks.bbox_mask(_to_device(_xyz(self.data)), min_x, min_y, min_z, max_x, max_y, max_z)
            self.data = self.data[keep.cpu().numpy()]
        debug_print(f"[DEBUG] Number of vertices after cropping: {len(self.data)}")
        status_print(f"After cropping, retained {len(self.data)} vertices.")
        return self.data

    # ------------------------------------------------------------------ host-only helpers (:233-354)
    @staticmethod
    def _dc_triplet(vertices):
        for prefix in ("", "scalar_", "scalar_scalar_"):
            names = [f"{prefix}f_dc_{i}" for i in range(3)]
            if names[0] in vertices.dtype.names:
                return np.column_stack([vertices[nm] for nm in names])
        return None

    @staticmethod
    def _compute_rgb_from_sh(vertices):
        """sRGB u8 from the SH DC term: clip(0.5 + C0*dc, 0, 1) ** (1/2.2) * 255."""
        dc = DataProcessor._dc_triplet(vertices)
        if dc is None:
            return None
        lin = np.clip(0.5 + dc * _SH_C0, 0.0, 1.0)
        return (np.power(lin, 1.0 / 2.2) * 255).astype(np.uint8)

    def add_rgb_from_sh(self):
        debug_print("[DEBUG] Executing 'add_rgb_from_sh' function...")
        names = self.data.dtype.names
        if "red" in names:
            return
        if "f_dc_0" not in names and "scalar_f_dc_0" not in names:
            debug_print("[DEBUG] No SH DC components found, cannot compute RGB.")
            return
        rgb = self._compute_rgb_from_sh(self.data)
        if rgb is None:
            return
        out = np.empty(len(self.data), dtype=self.data.dtype.descr + [("red", "u1"), ("green", "u1"), ("blue", "u1")])
        for nm in names:
            out[nm] = self.data[nm]
        out["red"], out["green"], out["blue"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
        self.data = out

    def cap_sh_degree(self, degree):
        """Zero the f_rest_* coefficients above `degree` (0 -> all 45, 1 -> from 9, 2 -> from 24)."""
        if degree is None or degree >= 3:
            return self.data
        first = {0: 0, 1: 9, 2: 24}.get(degree, 45)
        for i in range(first, 45):
            nm = f"f_rest_{i}"
            if nm in self.data.dtype.names:
                self.data[nm] = 0.0
        return self.data

    def apply_auto_bbox(self):
        """Report the tight bounding box of what is left (no change to the data)."""
        if len(self.data) == 0:
            status_print("Auto-BBox: No points remaining. Bounding box is undefined.")
            return
        lo = [np.min(self.data[a]) for a in "xyz"]
        hi = [np.max(self.data[a]) for a in "xyz"]
        status_print(f"Auto-BBox Applied: [{lo[0]:.4f}, {lo[1]:.4f}, {lo[2]:.4f}] to "
                     f"[{hi[0]:.4f}, {hi[1]:.4f}, {hi[2]:.4f}]")
