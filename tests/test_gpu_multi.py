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


def _lcg_start(dim, n):
    import numpy as np

    state = 12345
    out = np.empty((dim, n))
    for j in range(n):
        for r in range(dim):
            state = (state * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            out[r, j] = 4.0 * ((state >> 11) / 9007199254740992.0) - 2.0
    return out


def test_facade_devices(tmp_path):
    """Multi-GPU from the facades (SVGDOptions::Devices; reference: the Parallel flag, SVGD.hpp:49, 239-249): one process, one host
    thread per GPU inside the calls.  C++ program: one GPU against two, both against the oracle.  Python facade: F64 and TC32."""
    import numpy as np
    import oracle_binding as oracle

    ranks = _ranks()
    if ranks < 2:
        pytest.skip("needs at least two GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_facade_gpu import _build_and_run

    devs = [str(k) for k in range(min(ranks, 4))]
    out = _build_and_run("multi_device", tmp_path, src_dir=os.path.join("tests", "cpp"), args=devs)
    lines = [l for l in out.strip().splitlines() if not l.startswith("NCCL version")]  # (NCCL announces itself on stdout)
    dim, n, iters = 3, 700, 6
    assert lines[0] == "devices 1" and lines[dim + 1] == "devices %d" % len(devs)
    one = np.array([[float(t) for t in l.split()] for l in lines[1:dim + 1]])
    many = np.array([[float(t) for t in l.split()] for l in lines[dim + 2:2 * dim + 2]])
    mean = np.array([[0.5, -0.25, 1.0]])
    cov = np.array([[[0.9, 0.2, -0.1], [0.2, 0.8, 0.3], [-0.1, 0.3, 1.2]]])
    X0 = np.ascontiguousarray(_lcg_start(dim, n).T)
    ref = oracle.svgd_run(X0, iters, mean, cov, opt_kind=oracle.OPT_ADAM, lr=0.1)
    e1 = np.max(np.abs(one.T - ref)) / np.max(np.abs(ref))
    e2 = np.max(np.abs(many.T - ref)) / np.max(np.abs(ref))
    print("C++ facade: 1 GPU rel err %.3g, %d GPUs rel err %.3g" % (e1, len(devs), e2))
    assert e1 < 1e-9 and e2 < 1e-9

    import svgdcpp_b200 as sv

    for precision, tol in ((0, 1e-9), (1, 1e-3)):
        x0 = np.asfortranarray(_lcg_start(dim, n))
        model = sv.MultivariateNormal(mean[0], cov[0])
        svgd = sv.SVGD(dim, iters, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(dim, n, 0.1, 0.9, 0.999),
                       precision=precision, devices=[0, 1])
        assert svgd.NumDevices() == 2
        phi, a = svgd.ComputePhi()
        a_ref = oracle.rbf_median_scale(X0)
        svgd.Initialize()
        svgd.Run()
        svgd.close()
        err = np.sqrt(np.mean((x0.T - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
        print("Python facade, precision %d, 2 GPUs: scale rel err %.3g, trajectory rms rel err %.3g" % (precision, abs(a - a_ref) / a_ref, err))
        assert abs(a - a_ref) <= (1e-12 if precision == 0 else 1e-5) * a_ref
        assert err < tol
