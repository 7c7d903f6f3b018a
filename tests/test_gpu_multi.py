"""Row-sharded multi-GPU path (NCCL all-gather of the operands, all-reduced median counts): runs when the box has at least two
GPUs, with as many ranks as it has GPUs (up to 8); tests/multi_gpu_check.py does the work under torch.distributed.run.  The small
cases of that script leave ranks without any tile pair of the distance pass at 4 and 8 ranks (N = 301: two 256-row pairs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ranks():
    import torch

    return min(8, torch.cuda.device_count())


@pytest.mark.parametrize("precision", ["f64", "tc32"])
def test_multi_rank_parity(precision):
    ranks = _ranks()
    if ranks < 2:
        pytest.skip("needs at least two GPUs")
    for world in sorted({2, ranks}):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29531 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--precision", precision]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        sys.stdout.write(res.stdout[-4000:])
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
