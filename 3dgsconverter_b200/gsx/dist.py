"""Multi-GPU (one process per GPU, torch.distributed) versions of the hot path.

SOR (Taichi semantics) is global by construction: the bucket table is a hash of size N_global, so a
halo exchange cannot reproduce it.  Scheme: every rank holds a slab of the cloud;
  1. all-gather of the float32 xyz slabs (12 B/pt over NVLink),
  2. every rank builds the full hash grid (replicated, ~5 % of the query cost),
  3. rank r queries the r-th contiguous range of hash-sorted positions (neighbouring queries share
     buckets, so the range split keeps cache locality) and writes mean distances at the original indices
     of a zero-filled vector,
  4. ONE all-reduce(sum) of that vector -- every entry is written by exactly one rank, so x+0 is exact --
  5. the bit-exact NumPy-order mean/std and the threshold run replicated; each rank keeps its slab's mask.
The result is bit-identical to the single-GPU filter on the concatenated cloud.

K-Means (SOG chunks) shards by problem: chunks are independent, no data-path collective.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class _GsxOps:
    """Device ops used by the sharded drivers (replaceable in CPU/gloo tests)."""

    def build(self, xyz_all):
        from . import sor
        return sor.build_grid(xyz_all)

    def mean_dists_range(self, grid, k, hash_mode, out, qb, qe):
        from . import sor
        sor.mean_dists(grid, k, hash_mode, out=out, q_range=(qb, qe))

    def mask_from_means(self, means, threshold_factor):
        from . import sor
        return sor.threshold_mask(means, sor.mean_std(means), threshold_factor)


def _all_gather_rows(x: torch.Tensor, group=None):
    """All-gather a [n_local, C] tensor with possibly different n_local per rank."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out, sizes
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0), sizes


def query_range(n_total: int, rank: int, world: int):
    """Contiguous split of the hash-sorted query positions."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def sor_filter_sharded(xyz_local: torch.Tensor, k: int = 25, threshold_factor: float = 1.0,
                       hash_mode: str | None = None, group=None, return_means: bool = False, ops=None):
    """SOR keep-mask of this rank's slab, bit-identical to the single-GPU filter on the union cloud."""
    ops = ops or _GsxOps()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    xyz_all, sizes = _all_gather_rows(xyz_local, group)
    n = xyz_all.shape[0]
    grid = ops.build(xyz_all)
    qb, qe = query_range(n, rank, world)
    means = torch.zeros(n, dtype=torch.float32, device=xyz_all.device)
    ops.mean_dists_range(grid, k, hash_mode, means, qb, qe)
    dist.all_reduce(means, op=dist.ReduceOp.SUM, group=group)
    mask_all = ops.mask_from_means(means, threshold_factor)
    off = sum(sizes[:rank])
    sl = slice(off, off + sizes[rank])
    return (mask_all[sl], means[sl]) if return_means else mask_all[sl]


def kmeans_chunks_sharded(X_chunks, K: int, max_iter: int, inits, group=None, runner=None):
    """SOG shN schedule across ranks: chunk p goes to rank p % world; no collective on the data path.
    X_chunks / inits: lists (only the entries owned by this rank need to be real tensors).
    Returns {chunk index: (C, labels, counts)} for the chunks this rank owns."""
    from . import kmeans as gk
    runner = runner or gk.kmeans_lloyd
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    out = {}
    for p in range(len(X_chunks)):
        if p % world == rank:
            out[p] = runner(X_chunks[p], K, max_iter, inits[p])
    return out


class _GsxDensityOps:
    """Device ops of the sharded density filter (replaceable in CPU/gloo tests)."""

    def minmax(self, xyz):
        return torch.cat([xyz.min(dim=0).values, xyz.max(dim=0).values])

    def voxel_range(self, mm, voxel):
        from . import density
        return density.voxel_range(mm, voxel)

    def grid_count(self, xyz, voxel, q0, dim, grid):
        from . import density
        density.grid_count(xyz, voxel, q0, dim, grid)

    def grid_dense(self, grid, q0, dim, min_points, n_total):
        from . import density
        return density.grid_dense(grid, q0, dim, min_points, n_total)

    def member_mask(self, xyz, voxel, keep):
        from . import density
        return density.member_mask(xyz, voxel, keep)


def density_filter_sharded(xyz_local: torch.Tensor, voxel_size=1.0, threshold_percentage=0.32, sensitivity=None,
                           keep_multicluster=False, group=None, ops=None):
    """Density keep-mask of this rank's slab, identical to the single-GPU filter on the union cloud.
    all-reduce(min/max) of 6 floats -> global voxel box -> rank-local int32 histogram -> ONE all-reduce(sum)
    of the grid -> identical (tiny) host cluster selection on every rank -> local membership mask."""
    from . import density
    ops = ops or _GsxDensityOps()
    if sensitivity is not None:
        voxel_size, threshold_percentage = density.slider(sensitivity)
    n_local = torch.tensor([xyz_local.shape[0]], dtype=torch.int64, device=xyz_local.device)
    dist.all_reduce(n_local, group=group)
    n_total = int(n_local.item())
    mm = ops.minmax(xyz_local)
    lo, hi = mm[:3].clone(), mm[3:].clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    q0, dim = ops.voxel_range(torch.cat([lo, hi]).cpu().numpy(), voxel_size)
    ncell = int(dim[0]) * int(dim[1]) * int(dim[2])
    if ncell > density.GRID_CELL_LIMIT:
        # sparse far-flung cloud: replicate (all-gather) and run the single-GPU hash-table path on every rank
        xyz_all, sizes = _all_gather_rows(xyz_local, group)
        mask_all, info = density.density_filter(xyz_all, voxel_size, threshold_percentage, None, keep_multicluster)
        off = sum(sizes[: dist.get_rank(group)])
        return mask_all[off: off + xyz_local.shape[0]], info
    grid = torch.zeros(ncell, dtype=torch.int32, device=xyz_local.device)
    ops.grid_count(xyz_local, voxel_size, q0, dim, grid)
    dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
    min_points = int(n_total * (threshold_percentage / 100.0))  # data_processor.py:48 on the global count
    vox, cnt, n_unique = ops.grid_dense(grid, q0, dim, min_points, n_total)
    if len(vox) == 0:
        return torch.zeros(xyz_local.shape[0], dtype=torch.bool, device=xyz_local.device), dict(
            clusters=0, max_len=0, dense=0, voxels=n_unique)
    keep, n_kept, max_len = density.select_clusters(vox, keep_multicluster)
    mask = ops.member_mask(xyz_local, voxel_size, keep)
    return mask, dict(clusters=n_kept, max_len=max_len, dense=len(vox), voxels=n_unique)


# ------------------------------------------------------------------ distributed grid build (bucket-range ownership)
def _owner_bounds(n_global: int, world: int):
    """First bucket of every owner: rank o owns the buckets [ceil(o*N/G), ceil((o+1)*N/G))."""
    return [(o * n_global + world - 1) // world for o in range(world + 1)]


class _GsxBuildOps:
    """Device stages of the distributed grid build (replaceable in CPU/gloo tests)."""

    def cell_size(self, mm: np.ndarray, n_global: int) -> float:
        import ctypes as C
        from ._abi import lib
        return float(lib.gsx_sor_cell_size(mm.ctypes.data_as(C.POINTER(C.c_float)), n_global))

    def local_run(self, xyz_local, idx_base, n_global, world, bmin, cell):
        """A: stable partition of the slab by bucket owner.  Returns (pos4 [n,4] float32, cuts int64[world+1])."""
        import ctypes as C
        from . import sor
        from ._abi import lib, check
        from .sor import _ptr, _stream
        dev = xyz_local.device
        n_local = xyz_local.shape[0]
        ws_l = sor.workspace(max(n_local, 1), dev)
        pos4 = torch.empty((n_local, 4), dtype=torch.float32, device=dev)
        cuts = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        check(lib.gsx_sor_dist_local_run(_ptr(xyz_local), n_local, idx_base, n_global, world,
                                         bmin.ctypes.data_as(C.POINTER(C.c_float)), cell, _ptr(pos4), _ptr(cuts),
                                         _ptr(ws_l), ws_l.numel(), _stream()), "gsx_sor_dist_local_run")
        return pos4, cuts

    def merge_into(self, pos4_r, n_global, bmin, cell, out):
        """B: sort the received points of this rank's bucket range by (bucket, in-cell Morton) into `out`."""
        import ctypes as C
        from . import sor
        from ._abi import lib, check
        from .sor import _ptr, _stream
        m = pos4_r.shape[0]
        ws_m = sor.workspace(max(m, 1), pos4_r.device)
        check(lib.gsx_sor_dist_merge(_ptr(pos4_r), m, n_global, bmin.ctypes.data_as(C.POINTER(C.c_float)), cell,
                                     _ptr(out), _ptr(ws_m), ws_m.numel(), _stream()), "gsx_sor_dist_merge")

    def new_grid_storage(self, n_global, dev):
        """Workspace of the final grid and a [n_global,4] view of its sorted-position array (all-gather target)."""
        from . import sor
        from ._abi import lib
        ws = sor.workspace(n_global, dev)
        off = lib.gsx_sor_spos_offset(n_global)
        return ws, ws[off: off + n_global * 16].view(torch.float32).view(n_global, 4)

    def finish(self, ws, spos_full, n_global, bmin, cell):
        """C: table, boxes and bucket boxes from the globally sorted array."""
        import ctypes as C
        from . import sor
        from ._abi import lib, check
        from .sor import _ptr, _stream
        check(lib.gsx_sor_build_from_sorted(_ptr(spos_full), n_global, bmin.ctypes.data_as(C.POINTER(C.c_float)), cell,
                                            _ptr(ws), ws.numel(), _stream()), "gsx_sor_build_from_sorted")
        return sor.SorGrid(n_global, ws, bmin, cell)


def build_grid_distributed(xyz_local: torch.Tensor, group=None, ops=None):
    """Hash grid of the UNION cloud without replicating the sort: partition of the slab by bucket owner ->
    all-to-all -> owner-local sort -> all-gather of the sorted float4 segments -> table / boxes filled locally.
    Returns (grid over n_global points, sizes of the slabs)."""
    ops = ops or _GsxBuildOps()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = xyz_local.device
    n_local = xyz_local.shape[0]
    sizes_t = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes_t[rank] = n_local
    dist.all_reduce(sizes_t, group=group)
    sizes = [int(v) for v in sizes_t.tolist()]
    n_global = sum(sizes)
    idx_base = sum(sizes[:rank])
    # global bounding box -> cell size (gpu_ops.py:203-213 on the union cloud)
    lo = xyz_local.min(dim=0).values if n_local else torch.full((3,), float("inf"), device=dev)
    hi = xyz_local.max(dim=0).values if n_local else torch.full((3,), float("-inf"), device=dev)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    mm = torch.cat([lo, hi]).cpu().numpy().astype(np.float32)
    cell = ops.cell_size(mm, n_global)
    bmin = mm[:3].copy()
    # A. stable partition of the slab by bucket owner (one radix pass) + float4 gather
    pos4, cuts = ops.local_run(xyz_local, idx_base, n_global, world, bmin, cell)
    send = (cuts[1:] - cuts[:-1]).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    both = torch.stack([send, recv]).tolist()      # one host sync for both split lists
    send_l, recv_l = [int(v) for v in both[0]], [int(v) for v in both[1]]
    m = sum(recv_l)
    pos4_r = torch.empty((m, 4), dtype=torch.float32, device=dev)
    dist.all_to_all_single(pos4_r, pos4, recv_l, send_l, group=group)
    # B. owner-local sort, then all-gather of the segments in owner order = the globally hash-sorted array
    ws, spos_full = ops.new_grid_storage(n_global, dev)
    seg_sizes_t = torch.zeros(world, dtype=torch.int64, device=dev)
    seg_sizes_t[rank] = m
    dist.all_reduce(seg_sizes_t, group=group)
    seg_sizes = [int(v) for v in seg_sizes_t.tolist()]
    if len(set(seg_sizes)) == 1:
        seg = torch.empty((m, 4), dtype=torch.float32, device=dev)
        ops.merge_into(pos4_r, n_global, bmin, cell, seg)
        dist.all_gather_into_tensor(spos_full, seg, group=group)
    else:
        # ragged owners (N not a multiple of G): every rank sorts straight into its slot, then the slots are
        # broadcast one by one (sizes differ by at most one bucket)
        base = sum(seg_sizes[:rank])
        ops.merge_into(pos4_r, n_global, bmin, cell, spos_full[base: base + m])
        o = 0
        for r, sz in enumerate(seg_sizes):
            if sz:
                src = dist.get_global_rank(group, r) if group is not None else r
                dist.broadcast(spos_full[o: o + sz], src=src, group=group)
            o += sz
    # C. table, boxes, bucket boxes -- replicated, linear in n_global
    return ops.finish(ws, spos_full, n_global, bmin, cell), sizes


def sor_filter_auto(xyz_local, k=25, threshold_factor=1.0, hash_mode=None, group=None, return_means=False):
    """Picks the grid-build strategy by world size (replicated up to 3 ranks, distributed from 4)."""
    f = sor_filter_sharded_v2 if dist.get_world_size(group) >= 4 else sor_filter_sharded
    return f(xyz_local, k, threshold_factor, hash_mode, group=group, return_means=return_means)


def sor_filter_sharded_v2(xyz_local: torch.Tensor, k: int = 25, threshold_factor: float = 1.0,
                          hash_mode: str | None = None, group=None, return_means: bool = False, ops=None,
                          build_ops=None):
    """Like sor_filter_sharded, with the distributed grid build (no replicated sort, no all-gather of raw xyz)."""
    ops = ops or _GsxOps()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    grid, sizes = build_grid_distributed(xyz_local, group, ops=build_ops)
    n = sum(sizes)
    qb, qe = query_range(n, rank, world)
    means = torch.zeros(n, dtype=torch.float32, device=xyz_local.device)
    ops.mean_dists_range(grid, k, hash_mode, means, qb, qe)
    dist.all_reduce(means, op=dist.ReduceOp.SUM, group=group)
    mask_all = ops.mask_from_means(means, threshold_factor)
    off = sum(sizes[:rank])
    sl = slice(off, off + sizes[rank])
    return (mask_all[sl], means[sl]) if return_means else mask_all[sl]
