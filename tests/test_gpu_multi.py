"""Row-sharded multi-GPU path (NCCL all-gather of X and V, all-reduced median counts): runs only when the box has
at least two GPUs; tests/multi_gpu_check.py does the work under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["f64", "tc32"])
def test_two_rank_parity(precision):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--precision", precision]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
