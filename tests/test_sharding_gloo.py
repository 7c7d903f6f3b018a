"""World-size-2 `gloo` test (CPU) of the multi-GPU protocol of DESIGN.md "Multi-GPU": particle rows are sharded,
X and V are all-gathered, the median is found from all-reduced counts / histograms over a cyclic deal of the
symmetric tile pairs, and no reduction of phi is needed.  The arithmetic is the CPU oracle's; what is tested is
the decomposition the CUDA library implements (svgdcpp_b200/csrc/svgd_b200_api.cu: alloc_sharded,
launch_dist_pass, run_select, allgather_rows)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_plan(n, world, rank):
    """alloc_sharded(): contiguous blocks of ceil(n / world) rows, the last one short."""
    rpr = (n + world - 1) // world
    row0 = rpr * rank
    return rpr, row0, max(0, min(rpr, n - row0))


def _worker(rank, world, port, n, d, iters, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_binding as oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    A = rng.standard_normal((d, d))
    cov = (A @ A.T / d + 0.5 * np.eye(d))[None]
    mu = rng.standard_normal((1, d))
    X = 2.0 * rng.standard_normal((n, d))          # every rank starts from the same replica (svgdb_set_particles)
    rpr, row0, n_rows = shard_plan(n, world, rank)
    s1 = np.zeros((n_rows, d)); s2 = np.zeros((n_rows, d))   # optimizer state stays local to the row block
    T = (n + 63) // 64
    pairs = [(ti, tj) for ti in range(T) for tj in range(ti, T)]
    for it in range(iters):
        # --- median: symmetric tile pairs dealt cyclically to the ranks, counts all-reduced -----------------------
        mine = pairs[rank::world]
        vals = []
        for ti, tj in mine:
            Xi, Xj = X[ti * 64:(ti + 1) * 64], X[tj * 64:(tj + 1) * 64]
            D2 = np.maximum((Xi ** 2).sum(1)[:, None] + (Xj ** 2).sum(1)[None, :] - 2 * Xi @ Xj.T, 0.0)
            if ti == tj:
                np.fill_diagonal(D2, 0.0)
                vals.append(D2.ravel())
            else:
                vals.append(np.repeat(D2.ravel(), 2))      # weight 2: the transposed tile is never visited
        local = np.concatenate(vals) if vals else np.zeros(0)
        total, k_hi = n * n, (n * n) // 2
        keys = local.view(np.uint64)                           # IEEE bits of D2 >= 0 are order preserving

        def kth(k):                                            # radix select, one key bit per all-reduced count
            prefix = np.uint64(0)
            for bit in range(62, -1, -1):
                cand = prefix | (np.uint64(1) << np.uint64(bit))
                c = torch.tensor([float(np.sum(keys < cand))], dtype=torch.float64)
                dist.all_reduce(c)
                if c.item() <= k:
                    prefix = cand
            return np.array([prefix], dtype=np.uint64).view(np.float64)[0]

        med = 0.5 * (np.sqrt(kth(k_hi - 1)) + np.sqrt(kth(k_hi))) if total % 2 == 0 else np.sqrt(kth(k_hi))
        a = np.log(n) / med ** 2
        # --- local gradient rows, V, all-gather --------------------------------------------------------------------
        Xl = X[row0:row0 + n_rows]
        Gl = oracle.mvn_sum_logp_grad(Xl, mu, cov) if n_rows else np.zeros((0, d))
        Vl = np.zeros((rpr, d)); Vl[:n_rows] = Gl - 2 * a * Xl
        Vt = [torch.zeros(rpr, d, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(Vt, torch.from_numpy(Vl))
        V = torch.cat(Vt).numpy()[:n]
        # --- phi for the local rows against all columns; no reduction needed -------------------------------------
        D2 = ((Xl[:, None, :] - X[None, :, :]) ** 2).sum(-1)
        K = np.exp(-a * D2)
        phi = (K @ V + 2 * a * Xl * K.sum(1, keepdims=True)) / n
        s2 = 0.9 * s2 + 0.1 * phi
        s1 = 0.999 * s1 + 0.001 * phi ** 2
        step = 0.1 * (s2 / (1 - 0.9 ** (it + 1))) / (1e-8 + np.sqrt(s1 / (1 - 0.999 ** (it + 1))))
        Xn = np.zeros((rpr, d)); Xn[:n_rows] = Xl + step
        Xt = [torch.zeros(rpr, d, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(Xt, torch.from_numpy(Xn))
        X = torch.cat(Xt).numpy()[:n].copy()
    if rank == 0:
        np.save(out, X)
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d", [(130, 3), (257, 8)])
def test_row_sharded_protocol_matches_single_process_oracle(oracle, tmp_path, n, d):
    iters, world = 3, 2
    out = str(tmp_path / "x.npy")
    port = 29600 + (n % 50)
    mp.spawn(_worker, args=(world, port, n, d, iters, out), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    A = rng.standard_normal((d, d))
    cov = (A @ A.T / d + 0.5 * np.eye(d))[None]
    mu = rng.standard_normal((1, d))
    X0 = 2.0 * rng.standard_normal((n, d))
    ref = oracle.svgd_run(X0, iters, mu, cov, opt_kind=oracle.OPT_ADAM, lr=0.1)
    got = np.load(out)
    assert np.max(np.abs(got - ref)) < 1e-9 * np.max(np.abs(ref))


def _hessian_worker(rank, world, port, n, d, iters, out):
    """ScaleMethod::Hessian across ranks (hessian_scale_dev / prepare_and_phi_hessian): per-rank partial sums of -Hessian(log p)
    over the local rows, one all-reduce, A = R^T R, the scalar-bandwidth interaction with a = 1 on y = R x, g^ = R^-T g for the
    local rows against all columns, phi = R^T phi^."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_binding as oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    mus = 0.4 * rng.standard_normal((2, d))
    covs = np.stack([(lambda M: M @ M.T / d + 0.7 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(2)])
    X = 1.2 * rng.standard_normal((n, d))
    rpr, row0, n_rows = shard_plan(n, world, rank)
    s = np.zeros((n_rows, d))                                   # AdaGrad state of the local rows
    for it in range(iters):
        Xl = X[row0:row0 + n_rows]
        part = oracle.rbf_hessian_scale(Xl, mus, covs) * (2.0 * d * n_rows) if n_rows else np.zeros((d, d))
        H = torch.from_numpy(part.copy())
        dist.all_reduce(H)                                      # sum over ranks of sum_i -Hessian(log p)(x_i)
        A = H.numpy() / (2.0 * d * n)
        R = np.linalg.cholesky(A).T                             # A = R^T R, R upper triangular
        Y = X @ R.T                                             # every rank holds all particles
        Gh = oracle.mvn_sum_logp_grad(Xl, mus, covs) @ np.linalg.inv(R) if n_rows else np.zeros((0, d))
        Vl = np.zeros((rpr, d)); Vl[:n_rows] = Gh - 2.0 * Y[row0:row0 + n_rows]
        Vt = [torch.zeros(rpr, d, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(Vt, torch.from_numpy(Vl))
        V = torch.cat(Vt).numpy()[:n]
        Yl = Y[row0:row0 + n_rows]
        K = np.exp(-((Yl[:, None, :] - Y[None, :, :]) ** 2).sum(-1))
        phi = ((K @ V + 2.0 * Yl * K.sum(1, keepdims=True)) / n) @ R
        s = s + phi ** 2
        Xn = np.zeros((rpr, d)); Xn[:n_rows] = Xl + 0.1 * phi / (1e-8 + np.sqrt(s))
        Xt = [torch.zeros(rpr, d, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(Xt, torch.from_numpy(Xn))
        X = torch.cat(Xt).numpy()[:n].copy()
    if rank == 0:
        np.save(out, X)
    dist.destroy_process_group()


def test_hessian_scale_protocol_matches_single_process_oracle(oracle, tmp_path):
    n, d, iters, world = 131, 4, 3, 2
    out = str(tmp_path / "xh.npy")
    mp.spawn(_hessian_worker, args=(world, 29671, n, d, iters, out), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    mus = 0.4 * rng.standard_normal((2, d))
    covs = np.stack([(lambda M: M @ M.T / d + 0.7 * np.eye(d))(rng.standard_normal((d, d))) for _ in range(2)])
    X0 = 1.2 * rng.standard_normal((n, d))
    ref = oracle.svgd_run(X0, iters, mus, covs, opt_kind=oracle.OPT_ADAGRAD, lr=0.1, scale_method=oracle.SCALE_HESSIAN)
    got = np.load(out)
    assert np.max(np.abs(got - ref)) < 1e-9 * np.max(np.abs(ref))


def test_shard_plan_covers_every_row_once():
    for n in (1, 7, 128, 129, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            if world > n:
                continue
            seen = 0
            for r in range(world):
                rpr, row0, rows = shard_plan(n, world, r)
                assert row0 == seen or rows == 0
                seen += rows
            assert seen == n
