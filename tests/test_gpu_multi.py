"""Multi-GPU parity (peer-window / NCCL exchange along the reference's row/column groups, partitioned ingest):
tools/multi_gpu_check.py under torchrun on every GPU count the box offers.  Skipped on a single-GPU box; the CPU-side
plan is covered by tests/test_dist_cpu.py and tests/test_peer_protocol.py, profiles/r02_multi_gpu_check_p{2,4,8}.log record
the runs, and bench.py repeats a parity block at every N > 1 (`multi_gpu_parity`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_parity(n):
    if _ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + n), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MULTI_GPU_CHECK PASS") == n and "FAIL" not in out.stdout
