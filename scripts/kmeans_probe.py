"""K-Means timing probe (dev aid): python scripts/kmeans_probe.py [nprob] [rows] [D] [K] [iters]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))
import torch
from gsx import kmeans as gk

def main():
    a = [int(x) for x in sys.argv[1:]]
    nprob, rows, D, K, iters = (a + [8, 781_250, 45, 256, 2][len(a):])[:5]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(20260923)
    proto = torch.randn(1024, D, device=dev, generator=g) * 0.15
    idx = torch.randint(0, 1024, (nprob * rows,), device=dev, generator=g)
    X = proto[idx] + 0.03 * torch.randn(nprob * rows, D, device=dev, generator=g)
    offs = [p * rows for p in range(nprob + 1)]
    init = torch.stack([X[offs[p]:offs[p] + K] for p in range(nprob)])
    gk.kmeans_lloyd_batched(X, offs, K, 1, init)
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a_.record()
    gk.kmeans_lloyd_batched(X, offs, K, iters, init)
    b_.record(); torch.cuda.synchronize()
    ms = a_.elapsed_time(b_)
    fl = 3.0 * nprob * rows * K * D * iters
    print(f"nprob={nprob} rows={rows} D={D} K={K} iters={iters}: {ms:.2f} ms, {nprob*iters/(ms*1e-3):.1f} chunk-it/s, "
          f"{fl/(ms*1e-3)/1e12:.2f} T lane-instr/s (assign flops only)")

if __name__ == "__main__":
    main()
