#!/bin/bash
# like variant_probe.sh but i32wrap only and one process per variant over both clouds (dev aid)
cd "$(dirname "$0")/.."
for so in 3dgsconverter_b200/lib/variants/libgsx_*.so; do
  name=$(basename $so .so)
  GSX_LIB=$PWD/$so python - "$name" <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "3dgsconverter_b200")
import torch
from gsx import sor, synth
name = sys.argv[1]
out = [name]
for kind in ("mixed", "uniform"):
    x = torch.from_numpy(synth.xyz(10_000_000, kind)).cuda()
    ws = sor.workspace(x.shape[0], x.device)
    for mode in ("i32wrap", "i64"):
        ts = []
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); sor.sor_filter(x, 16, 2.0, hash_mode=mode, ws=ws); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out.append(f"{kind}/{mode} {min(ts[1:]):.2f}")
print("  ".join(out), flush=True)
PY
done
