"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/mgpu_worker.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_parity(n):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29540 + n), os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("MGPU_OK") == n
