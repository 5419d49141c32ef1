"""Small fixed workload for ncu: 8 chunks x 781250 x 45, K=256, `--iters` Lloyd iterations in `--mode`."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))
from gsx import kmeans as gk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="tensor")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--chunks", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
nprob, rows, D, K = a.chunks, 781_250, 45, 256
g = torch.Generator(device=dev).manual_seed(1)
proto = torch.randn(1024, D, device=dev, generator=g) * 0.15
X = proto[torch.randint(0, 1024, (nprob * rows,), device=dev, generator=g)] + 0.03 * torch.randn(nprob * rows, D, device=dev, generator=g)
offs = [p * rows for p in range(nprob + 1)]
init = torch.stack([X[offs[p]:offs[p] + K] for p in range(nprob)])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
gk.kmeans_lloyd_batched(X, offs, K, a.iters, init, assign=a.mode)
ev[1].record()
torch.cuda.synchronize()
print(f"mode={a.mode} chunks={nprob} iters={a.iters} ms={ev[0].elapsed_time(ev[1]):.3f}")
