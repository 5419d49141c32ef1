"""Upload / download time of a pageable 120 MB (C2 cloud) and 1.1 GB buffer through gsx_copy_h2d / gsx_copy_d2h for several
GSX_COPY_THREADS, next to plain cudaMemcpy from pageable and from pinned memory (development aid)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import sys, json, time
sys.path.insert(0, "."); sys.path.insert(0, "3dgsconverter_b200")
import numpy as np, torch
from gsx import hostcopy
res = {}
for mb in (120, 1100):
    a = np.random.default_rng(0).integers(0, 255, mb * 1000 * 1000, dtype=np.uint8)
    for name, up in (("staged", lambda: hostcopy.to_device(a, "cuda")), ("torch_pageable", lambda: torch.from_numpy(a).cuda())):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); t = up(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        res[f"h2d_{mb}MB_{name}_GBps"] = round(mb / 1e3 / min(ts[1:]), 1)
    for name, down in (("staged", lambda: hostcopy.to_host(t)), ("torch_pageable", lambda: t.cpu())):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); b = down(); ts.append(time.perf_counter() - t0)
        res[f"d2h_{mb}MB_{name}_GBps"] = round(mb / 1e3 / min(ts[1:]), 1)
    from gsx._abi import lib
    dst = np.zeros(a.nbytes, dtype=np.uint8)      # touched destination (what prefault_host buys the *_host entries)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lib.gsx_copy_d2h(dst.ctypes.data, t.data_ptr(), dst.nbytes, None); ts.append(time.perf_counter() - t0)
    res[f"d2h_{mb}MB_staged_touched_dst_GBps"] = round(mb / 1e3 / min(ts[1:]), 1)
    assert np.array_equal(dst, a)
    del dst
    p = torch.from_numpy(a).pin_memory()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); t = p.cuda(non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    res[f"h2d_{mb}MB_pinned_GBps"] = round(mb / 1e3 / min(ts[1:]), 1)
    del p, t
print(json.dumps(res))
'''
out = {}
for T, chunk_kb in ((8, 1024), (8, 4096), (4, 1024), (16, 1024), (8, 256)):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, GSX_COPY_THREADS=str(T), GSX_COPY_CHUNK_KB=str(chunk_kb)),
                       capture_output=True, text=True, timeout=300)
    key = f"threads={T},chunk_kb={chunk_kb}"
    out[key] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-300:]}
    print(key, json.dumps(out[key]), flush=True)
json.dump(out, open("gpurun_out/copy_threads_probe.json", "w"), indent=1)
