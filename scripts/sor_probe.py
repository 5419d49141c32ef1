"""Quick device-side timing probe for the SOR stages (development aid, not the bench)."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgsconverter_b200"))
import numpy as np, torch
from gsx import sor, synth

def ev_time(fn, reps=3):
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), ts

def main():
    dev = torch.device("cuda:0")
    sizes = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_000_000, 10_000_000]
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["uniform", "mixed"]
    for n in sizes:
        for kind in kinds:
            xyz = torch.from_numpy(synth.xyz(n, kind)).to(dev)
            ws = sor.workspace(n, dev)
            grid = sor.build_grid(xyz, ws)
            tb, _ = ev_time(lambda: sor.build_grid(xyz, ws))
            for mode in ("i32wrap", "i64"):
                out, st = sor.mean_dists(grid, 16, mode, want_stats=True)
                tq, all_ = ev_time(lambda: sor.mean_dists(grid, 16, mode))
                tf, _ = ev_time(lambda: sor.sor_filter(xyz, 16, 2.0, hash_mode=mode, ws=ws))
                print(json.dumps(dict(n=n, kind=kind, mode=mode, build_ms=round(tb, 3), knn_ms=round(tq, 3),
                                      filter_ms=round(tf, 3), msplats_s=round(n / tf / 1e3, 2),
                                      V_per_pt=round(st["visits"] / n, 1), scanned_per_pt=round(st["scanned"] / n, 1),
                                      boxes_per_pt=round(st["box_tests"] / n, 1), knn_all=[round(t, 2) for t in all_])),
                      flush=True)

if __name__ == "__main__":
    main()
