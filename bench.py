#!/usr/bin/env python
"""bench.py -- headline benchmark of the gsx hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gsx|reference] [--n POINTS] [--kind mixed]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Metric: Msplats/s of Statistical Outlier Removal, k=16, sigma=2.0 (BASELINE.json `metric`), on the
configs[1] workload: a 10 M-splat synthetic `mixed` cloud per GPU (SURVEY §8d generator), Taichi
semantics with the faithful int32-wrapping probe hash.  A step = one full pass of the filter
(min/max -> hash grid build -> K-nearest mean distances -> NumPy-order mean/std -> keep-mask) over the
batch, inputs resident in HBM.  N>1: weak scaling -- every rank holds a 10 M slab, the filter is the
GLOBAL one over the union cloud (gsx/dist.py: all-gather xyz, replicated grid, sharded queries, one
all-reduce of the mean distances), bit-identical to the single-GPU result.

Prints ONE JSON line (rank 0).  Extra objects: roofline (dominant kernel k_sor_knn), cpu_baseline (the
reference's CPU path timed on this host), e2e (through gsconverter.processing.gpu_ops.filter_sor_gpu with
host buffers), clocks, kmeans (secondary metric: K-Means chunk-iterations/s).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))

import numpy as np  # noqa: E402

K_SOR = 16
SIGMA = 2.0
L2_FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsx", choices=["gsx", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000, help="splats per GPU")
    ap.add_argument("--kind", default="mixed", choices=["mixed", "uniform", "clustered"])
    ap.add_argument("--hash", default="i32wrap", choices=["i32wrap", "i64"])
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="points of the CPU-baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / e2e / kmeans / i64 extras")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_sor(xyz_sample: np.ndarray, k: int, sigma: float):
    """The reference's CPU implementation of the path (data_processor.py:155-180): SciPy cKDTree, exact
    (k+1)-NN on cpu_count()-1 workers, mean/std threshold.  (Port: the reference is pure Python and is not
    present on the GPU box; oracle/sor.py restates those lines with the same SciPy calls.)"""
    import oracle
    t0 = time.perf_counter()
    md = oracle.sor_ckdtree_mean_dists(xyz_sample, k, workers=max(1, (os.cpu_count() or 2) - 1))
    mask = oracle.threshold_mask(md, sigma)
    return time.perf_counter() - t0, int(mask.sum())


def cpu_taichi_port_sor(xyz_sample: np.ndarray, k: int, sigma: float, mode: str):
    """Same algorithm as the GPU path (Taichi semantics) on all host cores: the C oracle."""
    import oracle
    t0 = time.perf_counter()
    md = oracle.sor_taichi_mean_dists(xyz_sample, k, mode)
    mask = oracle.threshold_mask(md, sigma)
    return time.perf_counter() - t0, int(mask.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gsx import synth
    ns = min(args.cpu_sample, args.n)
    xyz = synth.xyz(ns, args.kind)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_reference_sor(xyz, K_SOR, SIGMA)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = ns / t / 1e6
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "Msplats/s SOR k=16", "value": round(val, 4), "unit": "Msplats/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.n // 1_000_000}M-splat {args.kind} cloud per GPU, SOR k=16 sigma=2.0",
                   "sample": f"first {ns} points of the same cloud (bounded sample, whole filter per step)"},
        "cpu_baseline": {"value": round(val, 4), "unit": "Msplats/s", "cores": cores, "kind": "port",
                         "sample": f"{ns}-point prefix; reference CPU path = SciPy cKDTree k+1-NN "
                                   f"(data_processor.py:155-180) with workers=cpu_count()-1"},
        "e2e": {"value": round(val, 4), "unit": "Msplats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- gsx arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from gsx import sor, synth, _abi
    from gsx import dist as gd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    n = args.n
    blocks_per_rank = (n + synth.BLOCK - 1) // synth.BLOCK
    xyz_np = synth.xyz(n, args.kind, start_block=rank * blocks_per_rank)
    xyz = torch.from_numpy(xyz_np).to(dev)
    n_total = n * world
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    ws = sor.workspace(n_total, dev)
    means = torch.empty(n_total, dtype=torch.float32, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    knn_ms, build_ms = [], []

    def step(timed: bool):
        """One pass of the hot path; returns the keep-mask (device)."""
        flush.fill_(1)  # L2 flush between iterations (a 256 MiB write; < 0.3 % of a step)
        e0, e1, e2 = ev(), ev(), ev()
        if world == 1:
            e0.record()
            grid = sor.build_grid(xyz, ws)
            e1.record()
            sor.mean_dists(grid, K_SOR, args.hash, out=means)
            e2.record()
            mask = sor.threshold_mask(means, sor.mean_std(means), SIGMA)
        elif use_dist_build:
            # distributed grid build: local sort by global bucket key -> all-to-all by bucket owner -> owner sort
            # -> all-gather of the sorted float4 segments -> table/boxes filled locally (gsx/dist.py)
            e0.record()
            grid, sizes = gd.build_grid_distributed(xyz)
            e1.record()
            qb, qe = gd.query_range(n_total, rank, world)
            means.zero_()
            sor.mean_dists(grid, K_SOR, args.hash, out=means, q_range=(qb, qe))
            e2.record()
            dist.all_reduce(means)
            mask = sor.threshold_mask(means, sor.mean_std(means), SIGMA)[rank * n:(rank + 1) * n]
        else:
            xyz_all, sizes = gd._all_gather_rows(xyz)
            e0.record()
            grid = sor.build_grid(xyz_all, ws)
            e1.record()
            qb, qe = gd.query_range(n_total, rank, world)
            means.zero_()
            sor.mean_dists(grid, K_SOR, args.hash, out=means, q_range=(qb, qe))
            e2.record()
            dist.all_reduce(means)
            mask = sor.threshold_mask(means, sor.mean_std(means), SIGMA)[rank * n:(rank + 1) * n]
        if timed:
            step.events.append((e0, e1, e2))
        return mask

    step.events = []
    # replicated grid build (all-gather raw xyz, every rank sorts the union cloud) is faster up to 2-3 ranks;
    # from 4 ranks on the distributed build wins (measured: N=2 3.1 vs 3.9 ms, N=4 6.1 vs 5.1 ms).  Both are
    # bit-identical to the single-GPU filter (tests/test_multigpu_nccl.py).
    env_db = os.environ.get("GSX_DIST_BUILD", "auto")
    use_dist_build = world >= 4 if env_db == "auto" else env_db == "1"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        mask = step(False)
    barrier()
    launches0 = _abi.lib.gsx_kernel_launches()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t_start, t_end = ev(), ev()
    barrier()
    t_start.record()
    for _ in range(args.steps):
        mask = step(True)
    t_end.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = _abi.lib.gsx_kernel_launches() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    for e0, e1, e2 in step.events:
        build_ms.append(e0.elapsed_time(e1))
        knn_ms.append(e1.elapsed_time(e2))
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n_total / (ms_per_step * 1e-3) / 1e6
    kept = int(mask.sum().item())

    # ---- roofline of the dominant kernel (k_sor_knn): gather-model algorithmic bytes of OUR algorithm
    # B = queries*(16 own float4 + 27*32 bucket entries {start,end,box} + 4 result) + 16*candidates scanned
    #     + 32*chunk/super boxes tested,
    # counted exactly by the instrumented build of the same kernel.
    grid = sor.build_grid(xyz, ws) if world == 1 else gd.build_grid_distributed(xyz)[0]  # same grid either way
    qr = gd.query_range(n_total, rank, world)
    _, st = sor.mean_dists(grid, K_SOR, args.hash, out=means, want_stats=True, q_range=qr)
    alg_bytes = st["queries"] * (16 + 27 * 32 + 4) + 16 * st["scanned"] + 32 * st["box_tests"]
    ref_model_bytes = st["queries"] * (16 + 27 * 8 + 4) + 16 * st["visits"]
    knn_avg_ms = float(np.mean(knn_ms))
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (knn_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_sor_knn", "achieved": round(achieved, 1), "peak": peak,
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                "kernel_ms": round(knn_avg_ms, 3), "kernel_share_of_step": round(knn_avg_ms / ms_per_step, 3),
                "algorithmic_bytes_per_launch": int(alg_bytes),
                "per_query": {"ref_visits_V": round(st["visits"] / st["queries"], 1),
                              "scanned": round(st["scanned"] / st["queries"], 1),
                              "box_tests": round(st["box_tests"] / st["queries"], 1)},
                "reference_gather_model_GBps": round(ref_model_bytes / (knn_avg_ms * 1e-3) / 1e9, 1),
                "note": "achieved counts the bytes our pruned search gathers (L1/L2-served, DRAM traffic is far "
                        "lower); the reference's un-pruned gather model (16 B x V visits) would read "
                        "reference_gather_model_GBps"}
    prof = ROOT / "profiles" / "r01_knn_traffic.json"
    if prof.exists() and world == 1 and n == 10_000_000 and args.kind == "mixed" and args.hash == "i32wrap":
        try:
            roofline["traffic"] = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            pass

    line = {
        "metric": "Msplats/s SOR k=16", "value": round(value, 3), "unit": "Msplats/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n // 1_000_000}M-splat {args.kind} cloud per GPU (SURVEY 8d generator), SOR k=16 "
                               f"sigma=2.0, Taichi semantics, probe hash {args.hash}; global filter over the union "
                               f"cloud of {n_total} splats",
                   "splats_per_gpu": n, "k": K_SOR, "sigma": SIGMA, "hash_mode": args.hash,
                   "l2": "256 MiB flush write before every step + working set (~0.6 GB/step) larger than L2",
                   "kept": kept, "parallelism": (f"dp{world}: distributed grid build (all-to-all by bucket owner + "
                                                 "all-gather of sorted float4), sharded queries, one all-reduce"
                                                 if use_dist_build else
                                                 f"dp{world}: all-gather xyz, replicated grid, sharded queries, "
                                                 "one all-reduce") if world > 1 else "single GPU"},
        "stage_ms": {"build": round(float(np.mean(build_ms)), 3), "knn": round(knn_avg_ms, 3)},
        "gpu_launches": int(launches), "roofline": roofline,
    }
    if clk:
        line["clocks"] = clk

    # ---- extras on rank 0 at N=1: e2e through the plugin API, CPU baseline, K-Means secondary metric
    if world == 1 and not args.no_extras:
        line["e2e"] = measure_e2e(xyz_np, args)
        line["cpu_baseline"] = measure_cpu_baseline(xyz_np, args)
        line["other_modes"] = measure_other(xyz, ws, means, args)
        line["pipeline"] = measure_pipeline(xyz_np, args)
        line["kmeans"] = measure_kmeans(dev)
    elif world > 1:
        line["e2e"] = measure_e2e_sharded(xyz_np, args, dev, rank, world)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(xyz_np, args):
    """Same metric through the reference-facing plugin call with HOST buffers:
    gsconverter.processing.gpu_ops.filter_sor_gpu(np.ndarray) -> np.ndarray[bool];
    H2D of the xyz (pinned) and D2H of the mask are inside the timed region."""
    import torch
    os.environ["GSX_SOR_HASH"] = args.hash
    from gsconverter.processing import gpu_ops
    n = len(xyz_np)
    pinned = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    host = pinned.numpy()
    host[:] = xyz_np
    for _ in range(2):
        m = gpu_ops.filter_sor_gpu(host, K_SOR, SIGMA)
    torch.cuda.synchronize()
    reps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(reps):
        m = gpu_ops.filter_sor_gpu(host, K_SOR, SIGMA)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(n / dt / 1e6, 3), "unit": "Msplats/s", "ms_per_step": round(dt * 1e3, 3),
            "h2d_bytes_per_step": int(n * 12), "d2h_bytes_per_step": int(n),
            "api": "gsconverter.processing.gpu_ops.filter_sor_gpu(numpy[N,3]) -> numpy bool[N]", "kept": int(m.sum())}


def measure_e2e_sharded(xyz_np, args, dev, rank, world):
    """N>1: host slab -> device -> sharded global filter -> host mask, per rank; max over ranks."""
    import torch
    import torch.distributed as dist
    from gsx import dist as gd
    n = len(xyz_np)
    pinned = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    pinned.numpy()[:] = xyz_np
    out = torch.empty(n, dtype=torch.bool).pin_memory()

    def once():
        x = pinned.to(dev, non_blocking=True)
        env_db = os.environ.get("GSX_DIST_BUILD", "auto")
        dist_build = world >= 4 if env_db == "auto" else env_db == "1"
        mask = (gd.sor_filter_sharded_v2 if dist_build else gd.sor_filter_sharded)(x, K_SOR, SIGMA, args.hash)
        out.copy_(mask, non_blocking=True)
        torch.cuda.synchronize()
    once()
    dist.barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {"value": round(n * world / dt / 1e6, 3), "unit": "Msplats/s", "ms_per_step": round(dt * 1e3, 3),
            "h2d_bytes_per_step": int(n * 12), "d2h_bytes_per_step": int(n),
            "api": "gsx.dist.sor_filter_sharded(host slab per rank)"}


def measure_cpu_baseline(xyz_np, args):
    ns = min(args.cpu_sample, len(xyz_np))
    sample = np.ascontiguousarray(xyz_np[:ns])
    cores = os.cpu_count() or 1
    dt, kept = cpu_reference_sor(sample, K_SOR, SIGMA)
    dt2, kept2 = cpu_taichi_port_sor(sample, K_SOR, SIGMA, args.hash)
    return {"value": round(ns / dt / 1e6, 4), "unit": "Msplats/s", "cores": cores, "kind": "port",
            "sample": f"first {ns} points of the bench cloud, one pass; reference CPU path = SciPy cKDTree "
                      f"(k+1)-NN on cpu_count()-1={max(1, cores - 1)} workers + mean/std mask "
                      f"(data_processor.py:155-180)",
            "seconds": round(dt, 3),
            "taichi_semantics_port": {"value": round(ns / dt2 / 1e6, 4), "unit": "Msplats/s", "seconds": round(dt2, 3),
                                      "what": "C/OpenMP oracle of the Taichi kernel (same results as the GPU path), "
                                              "all host cores, same sample"}}


def measure_other(xyz, ws, means, args):
    """Whole-filter throughput for the other probe-hash mode and the uniform cloud (device-resident)."""
    import torch
    from gsx import sor, synth
    out = {}

    def timed(x, mode):
        for _ in range(2):
            sor.sor_filter(x, K_SOR, SIGMA, hash_mode=mode, ws=ws)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            sor.sor_filter(x, K_SOR, SIGMA, hash_mode=mode, ws=ws)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 5

    other = "i64" if args.hash == "i32wrap" else "i32wrap"
    ms = timed(xyz, other)
    out[f"{args.kind}_{other}"] = {"ms": round(ms, 3), "msplats_s": round(xyz.shape[0] / ms / 1e3, 2)}
    if args.kind != "uniform":
        xu = torch.from_numpy(synth.xyz(xyz.shape[0], "uniform")).to(xyz.device)
        for mode in ("i32wrap", "i64"):
            ms = timed(xu, mode)
            out[f"uniform_{mode}"] = {"ms": round(ms, 3), "msplats_s": round(xu.shape[0] / ms / 1e3, 2)}
    return out


def measure_pipeline(xyz_np, args):
    """BASELINE configs[1] chained (converter.py:194-236 order): bbox -> alpha(5) -> density(0.5, multicluster)
    -> SOR k=16 on the 10 M cloud through gsx.pipeline.FilterChain: pinned host xyz+opacity in, surviving row
    indices out (the device-resident working set; host records are gathered once by the caller)."""
    import torch
    from gsx import synth
    from gsx.pipeline import FilterChain
    n = len(xyz_np)
    op_np = synth.attributes(n, 0)["opacity"] if False else np.random.default_rng(1).normal(0, 2, n).astype(np.float32)
    px = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    px.numpy()[:] = xyz_np
    po = torch.empty(n, dtype=torch.float32).pin_memory()
    po.numpy()[:] = op_np

    def once():
        ch = FilterChain(px, po)
        c0 = ch.crop_by_bbox(-11, -11, -11, 11, 11, 11)
        c1 = ch.alpha(5)
        c2, _ = ch.density(sensitivity=0.5, keep_multicluster=True)
        c3 = ch.sor(K_SOR, SIGMA, hash_mode=args.hash)
        idx = ch.indices()
        return (c0, c1, c2, c3, len(idx))
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        counts = once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"what": "bbox -> alpha(5) -> density(0.5, multicluster) -> SOR k=16, host columns in, surviving indices out",
            "ms": round(dt * 1e3, 2), "msplats_s": round(n / dt / 1e6, 2), "survivors_per_stage": list(counts[:4]),
            "h2d_bytes": int(n * 16), "d2h_bytes": int(counts[4] * 4)}


def measure_kmeans(dev):
    """Secondary metric of BASELINE.json: K-Means iterations/s on the SOG shN schedule (sog.py:527-549),
    64 chunks x 781 250 x 45, K=256 (the 50 M-splat C3 config, 9 GB of SH rows), 2 Lloyd iterations."""
    import torch
    from gsx import kmeans as gk
    nprob, rows, D, K, iters = 64, 781_250, 45, 256, 2
    g = torch.Generator(device=dev).manual_seed(20260923)
    proto = torch.randn(1024, D, device=dev, generator=g) * 0.15
    X = torch.empty((nprob * rows, D), dtype=torch.float32, device=dev)
    for p in range(nprob):  # chunk-wise generation keeps the temporaries small
        idx = torch.randint(0, 1024, (rows,), device=dev, generator=g)
        X[p * rows:(p + 1) * rows] = proto[idx] + 0.03 * torch.randn(rows, D, device=dev, generator=g)
    offs = [p * rows for p in range(nprob + 1)]
    init = torch.stack([X[offs[p]:offs[p] + K] for p in range(nprob)])
    gk.kmeans_lloyd_batched(X, offs, K, 1, init)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    gk.kmeans_lloyd_batched(X, offs, K, iters, init)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    chunk_iters = nprob * iters
    flops = 3.0 * nprob * rows * K * D * iters
    out = {"metric": "K-Means chunk-iterations/s (781250x45, K=256)", "value": round(chunk_iters / (ms * 1e-3), 2),
           "ms_total": round(ms, 2), "chunks": nprob, "iters": iters,
           "fp32_lane_instr_per_s_T": round(flops / (ms * 1e-3) / 1e12, 2),
           "fp32_no_fma_peak_T": 37.2, "frac_of_fp32_peak": round(flops / (ms * 1e-3) / 1e12 / 37.2, 3),
           "mpoint_iters_per_s": round(nprob * rows * iters / (ms * 1e-3) / 1e6, 1)}
    # the exact fma pre-filter (bit-identical labels, off by default in round 1): timed last, best effort
    try:
        gk.set_prefilter(True)
        gk.kmeans_lloyd_batched(X, offs, K, 1, init)
        torch.cuda.synchronize()
        a.record()
        gk.kmeans_lloyd_batched(X, offs, K, iters, init)
        b.record()
        torch.cuda.synchronize()
        ms2 = a.elapsed_time(b)
        out["with_exact_prefilter"] = {"value": round(chunk_iters / (ms2 * 1e-3), 2), "ms_total": round(ms2, 2)}
    except Exception as e:  # noqa: BLE001
        out["with_exact_prefilter"] = {"error": str(e)[:200]}
    finally:
        gk.set_prefilter(False)
    # the reference's CPU path for the same call (gpu_ops.py:48-52: scikit-learn MiniBatchKMeans, unseeded --
    # a different algorithm whose "iterations" are mini-batch passes): one 781 250 x 45 chunk, K=256, max_iter=10
    try:
        from sklearn.cluster import MiniBatchKMeans
        xc = X[:rows].cpu().numpy()
        t0 = time.perf_counter()
        MiniBatchKMeans(n_clusters=K, max_iter=10, batch_size=min(4096 * 4, len(xc)), n_init="auto",
                        compute_labels=True).fit(xc)
        dt = time.perf_counter() - t0
        out["cpu_reference_sklearn"] = {"seconds_per_chunk_max_iter10": round(dt, 2),
                                        "chunk_fits_per_s": round(1.0 / dt, 3), "cores": os.cpu_count(),
                                        "note": "the reference's fallback when Taichi is absent; one full chunk"}
    except Exception as e:  # noqa: BLE001
        out["cpu_reference_sklearn"] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    main()
