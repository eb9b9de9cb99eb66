"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the row-sharded NCCL path must
reproduce the single-GPU solve bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_solve_is_bit_identical(torch_cuda):
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), "32,64,128,256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-3000:]); sys.stderr.write(out.stderr[-3000:])
    assert out.returncode == 0
    lines = [l for l in out.stdout.splitlines() if l.startswith("DIST ")]
    assert len(lines) == 4 and all(l.endswith("OK") for l in lines), lines
