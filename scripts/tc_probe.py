"""GPU probe of the tcgen05 K-Means assign: descriptor variants, score error vs float64, label parity and timing.
Writes gpurun_out/tc_probe.json.  (Diagnostics only -- the asserted checks live in tests/test_kmeans_tc_gpu.py.)"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))
from gsx import kmeans as gk  # noqa: E402

dev = torch.device("cuda:0")
out = {}
rng = np.random.default_rng(0)


def tf32(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


for (D, K) in ((45, 256), (45, 64), (24, 100), (9, 16)):
    X = rng.normal(0, 0.15, (128, D)).astype(np.float32)
    C = rng.normal(0, 0.15, (K, D)).astype(np.float32)
    exact = X.astype(np.float64) @ C.astype(np.float64).T - 0.5 * (C.astype(np.float64) ** 2).sum(1)[None]
    trunc = tf32(X).astype(np.float64) @ tf32(C).astype(np.float64).T - 0.5 * (C.astype(np.float64) ** 2).sum(1)[None]
    for variant in (0, 1):
        try:
            S = gk.tc_debug_scores(torch.from_numpy(X).to(dev), torch.from_numpy(C).to(dev), variant).cpu().numpy()[:, :K]
            out[f"scores_D{D}_K{K}_v{variant}"] = {
                "max_abs_err_vs_exact": float(np.abs(S - exact).max()),
                "max_abs_err_vs_tf32_trunc_inputs": float(np.abs(S - trunc).max()),
                "score_scale": float(np.abs(exact).max()),
                "sample": [float(S[0, 0]), float(exact[0, 0]), float(S[5, 7]), float(exact[5, 7]), float(S[127, K - 1]), float(exact[127, K - 1])]}
        except Exception as e:  # noqa: BLE001
            out[f"scores_D{D}_K{K}_v{variant}"] = {"error": str(e)[:300]}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "tc_probe.json").write_text(json.dumps(out, indent=1))

# label parity tensor vs strict on a medium problem, and timing on 8 chunks of the C3 shape
try:
    from gsx import synth
    Xn = np.ascontiguousarray(synth.attributes(200_000)["f_rest"])
    Xd = torch.from_numpy(Xn).to(dev)
    init = Xd[:256].clone()
    res = {}
    for mode in ("strict", "tensor"):
        Cc, L, cnt = gk.kmeans_lloyd(Xd, 256, 3, init, assign=mode)
        res[mode] = (Cc.cpu().numpy(), L.cpu().numpy())
    out["parity_200k"] = {"labels_equal": bool(np.array_equal(res["strict"][1], res["tensor"][1])),
                          "n_label_diff": int((res["strict"][1] != res["tensor"][1]).sum()),
                          "centroids_equal": bool(np.array_equal(res["strict"][0].view(np.uint32), res["tensor"][0].view(np.uint32)))}
    (ROOT / "gpurun_out" / "tc_probe.json").write_text(json.dumps(out, indent=1))
    nprob, rows, D, K = 8, 781_250, 45, 256
    g = torch.Generator(device=dev).manual_seed(1)
    proto = torch.randn(1024, D, device=dev, generator=g) * 0.15
    X = proto[torch.randint(0, 1024, (nprob * rows,), device=dev, generator=g)] + 0.03 * torch.randn(nprob * rows, D, device=dev, generator=g)
    offs = [p * rows for p in range(nprob + 1)]
    init = torch.stack([X[offs[p]:offs[p] + K] for p in range(nprob)])
    for mode in ("strict", "fma", "tensor"):
        gk.kmeans_lloyd_batched(X, offs, K, 1, init, assign=mode)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = gk.kmeans_lloyd_batched(X, offs, K, 3, init, assign=mode, want_stats=True)
        b.record()
        torch.cuda.synchronize()
        out[f"time_8chunks_3it_{mode}"] = {"ms": a.elapsed_time(b), "stats": r[3]}
        res[mode] = r[1].cpu().numpy()
    out["parity_8chunks"] = {"tensor_vs_strict_labels_equal": bool(np.array_equal(res["strict"], res["tensor"])),
                             "fma_vs_strict_labels_equal": bool(np.array_equal(res["strict"], res["fma"]))}
except Exception as e:  # noqa: BLE001
    out["parity_error"] = str(e)[:500]
(ROOT / "gpurun_out" / "tc_probe.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
