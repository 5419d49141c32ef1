"""CPU: csrc/gsx_hostcopy.cu (worker pool, chunk ring, D2H pipeline, prefault) compiled HOST-ONLY against a mock CUDA
runtime and run under ThreadSanitizer.  The mock's streams are real FIFO queues drained by background threads with a delay
per operation (tests/mock_cuda/cuda_runtime.h), so an "async" copy happens later than its enqueue: refilling a pinned
chunk before the event of its previous DMA, or handing a chunk to the caller before its DMA has landed, corrupts the data
(checked once by deleting the slot-reuse wait: the harness fails at 64 MiB).  Covered: byte-exact round trips around the
chunk / threshold boundaries, the "source fully read on return" contract, repeated calls on the parked worker threads,
several caller threads at once, no data race.  Real hardware ordering is covered by tests/test_hostcopy_gpu.py."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
MOCK = ROOT / "tests" / "mock_cuda"
CSRC = ROOT / "3dgsconverter_b200" / "csrc"


@pytest.mark.parametrize("threads", ["1", "6"])
def test_hostcopy_against_mock_runtime_under_tsan(tmp_path, threads):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "harness"
    base = [gxx, "-std=c++17", "-O1", "-g", f"-I{MOCK}", f"-I{CSRC}", "-x", "c++", str(MOCK / "hostcopy_harness.cpp"),
            "-o", str(exe), "-pthread"]
    r = subprocess.run(base[:4] + ["-fsanitize=thread"] + base[4:], capture_output=True, text=True)
    if r.returncode != 0:       # no libtsan in this toolchain: still run the functional part
        r = subprocess.run(base, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, GSX_COPY_THREADS=threads))
    assert run.returncode == 0 and "hostcopy mock harness OK" in run.stdout, (run.stdout + run.stderr)[-3000:]
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-3000:]
