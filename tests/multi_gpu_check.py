"""Multi-GPU parity check, launched with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--precision f64|tc32]

Each rank owns a contiguous block of particle rows (svgdb_comm_init); X and V are all-gathered with NCCL every
step and the median counts / select histograms are all-reduced.  Rank 0 compares kernel scale, phi and the
trajectory with the single-process CPU oracle.  Exit code 0 = parity.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="f64")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from svgdcpp_b200 import _capi

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _capi.load()
    prec = _capi.PRECISION_TC32 if args.precision == "tc32" else _capi.PRECISION_F64
    dp = C.POINTER(C.c_double)
    failures = 0
    for n, d, iters in [(301, 5, 6), (1000, 64, 5), (1500, 33, 4), (4100, 64, 2), (600, 160, 3)]:
        rng = np.random.default_rng(n)
        A = rng.standard_normal((d, d))
        cov = np.ascontiguousarray(A @ A.T / d + 0.5 * np.eye(d))
        mu = np.ascontiguousarray(rng.standard_normal(d))
        X0 = np.ascontiguousarray(2.0 * rng.standard_normal((n, d)))
        ctx = C.c_void_p()

        def check(rc):
            if rc != 0:
                raise RuntimeError(lib.svgdb_last_error(ctx).decode())

        check(lib.svgdb_create(C.byref(ctx), local, n, d, prec))
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            assert lib.svgdb_nccl_unique_id(uid.ctypes.data_as(C.c_void_p), 128) == 0
        t = torch.from_numpy(uid).cuda()
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
        check(lib.svgdb_comm_init(ctx, world, rank, uid.ctypes.data_as(C.c_void_p), 128))
        check(lib.svgdb_set_model_mvn(ctx, mu.ctypes.data_as(dp), cov.ctypes.data_as(dp)))
        check(lib.svgdb_set_kernel_rbf(ctx, _capi.SCALE_MEDIAN, 0.0))
        check(lib.svgdb_set_optimizer(ctx, _capi.OPT_ADAM, 0.1, 0.9, 0.999, 1e-8))
        check(lib.svgdb_set_particles(ctx, X0.ctypes.data_as(dp)))
        check(lib.svgdb_initialize(ctx))
        phi = np.empty_like(X0)
        a = C.c_double(0.0)
        check(lib.svgdb_compute_phi(ctx, phi.ctypes.data_as(dp), C.byref(a)))
        check(lib.svgdb_step(ctx, iters))
        X = np.empty_like(X0)
        check(lib.svgdb_get_particles(ctx, X.ctypes.data_as(dp)))
        s1 = np.empty_like(X0)
        cnt = C.c_uint64(0)
        check(lib.svgdb_get_opt_state(ctx, s1.ctypes.data_as(dp), None, C.byref(cnt)))
        # two more iterations through the host-buffer call, every rank moving only its own rows (svgdb_step_host under sharding)
        r0, nr = C.c_int64(0), C.c_int64(0)
        check(lib.svgdb_local_rows(ctx, C.byref(r0), C.byref(nr)))
        mine = X[r0.value:r0.value + nr.value].copy()  # (a contiguous slice would alias X)
        check(lib.svgdb_step_host(ctx, mine.ctypes.data_as(dp), mine.ctypes.data_as(dp), 2))
        X2 = np.empty_like(X0)
        check(lib.svgdb_get_particles(ctx, X2.ctypes.data_as(dp)))
        own_rows_ok = bool(np.array_equal(mine, X2[r0.value:r0.value + nr.value]))
        lib.svgdb_destroy(ctx)
        if rank == 0:
            import oracle_binding as oracle

            a_ref = oracle.rbf_median_scale(X0)
            G_ref = oracle.mvn_sum_logp_grad(X0, mu[None], cov[None], lse=True)
            phi_ref = oracle.phi(X0, G_ref, a_ref)
            X_ref = oracle.svgd_run(X0, iters, mu[None], cov[None], opt_kind=oracle.OPT_ADAM, lr=0.1, lse=True)
            e_a = abs(a.value - a_ref) / a_ref
            e_phi = np.max(np.abs(phi - phi_ref)) / np.max(np.abs(phi_ref))
            e_x = np.sqrt(np.mean((X - X_ref) ** 2)) / np.sqrt(np.mean(X_ref ** 2))
            tol = (1e-12, 1e-11, 1e-9) if prec == _capi.PRECISION_F64 else (1e-5, 2e-4, 1e-3)
            X2_ref = oracle.svgd_run(X0, iters + 2, mu[None], cov[None], opt_kind=oracle.OPT_ADAM, lr=0.1, lse=True)
            e_x2 = np.sqrt(np.mean((X2 - X2_ref) ** 2)) / np.sqrt(np.mean(X2_ref ** 2))
            ok = e_a < tol[0] and e_phi < tol[1] and e_x < tol[2] and e_x2 < tol[2] and cnt.value == iters and np.all(np.isfinite(s1))
            print("world=%d %s n=%d d=%d: a err %.2e, phi err %.2e, trajectory rms err %.2e, after 2 more steps through svgdb_step_host %.2e -> %s"
                  % (world, args.precision, n, d, e_a, e_phi, e_x, e_x2, "OK" if ok else "FAIL"), flush=True)
            failures += 0 if ok else 1
        if not own_rows_ok:
            print("rank %d: svgdb_step_host returned rows that differ from the device's copy" % rank, flush=True)
            failures += 1
    if prec == _capi.PRECISION_F64:  # ScaleMethod::Hessian: per-rank Hessian partial sums, one all-reduce
        n, d, iters = 700, 12, 4
        rng = np.random.default_rng(7)
        A = rng.standard_normal((d, d))
        cov = np.ascontiguousarray(A @ A.T / d + 0.5 * np.eye(d))
        mu = np.ascontiguousarray(rng.standard_normal(d))
        X0 = np.ascontiguousarray(2.0 * rng.standard_normal((n, d)))
        ctx = C.c_void_p()

        def check(rc):
            if rc != 0:
                raise RuntimeError(lib.svgdb_last_error(ctx).decode())

        check(lib.svgdb_create(C.byref(ctx), local, n, d, prec))
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            assert lib.svgdb_nccl_unique_id(uid.ctypes.data_as(C.c_void_p), 128) == 0
        t = torch.from_numpy(uid).cuda()
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
        check(lib.svgdb_comm_init(ctx, world, rank, uid.ctypes.data_as(C.c_void_p), 128))
        check(lib.svgdb_set_model_mvn(ctx, mu.ctypes.data_as(dp), cov.ctypes.data_as(dp)))
        check(lib.svgdb_set_kernel_rbf(ctx, _capi.SCALE_HESSIAN, 0.0))
        check(lib.svgdb_set_optimizer(ctx, _capi.OPT_ADAM, 0.1, 0.9, 0.999, 1e-8))
        check(lib.svgdb_set_particles(ctx, X0.ctypes.data_as(dp)))
        check(lib.svgdb_initialize(ctx))
        phi = np.empty_like(X0)
        a = C.c_double(0.0)
        check(lib.svgdb_compute_phi(ctx, phi.ctypes.data_as(dp), C.byref(a)))
        Amat = np.empty((d, d))
        check(lib.svgdb_get_scale_matrix(ctx, Amat.ctypes.data_as(dp)))
        check(lib.svgdb_step(ctx, iters))
        X = np.empty_like(X0)
        check(lib.svgdb_get_particles(ctx, X.ctypes.data_as(dp)))
        lib.svgdb_destroy(ctx)
        if rank == 0:
            import oracle_binding as oracle

            A_ref = oracle.rbf_hessian_scale(X0, mu[None], cov[None])
            phi_ref = oracle.phi_matrix(X0, oracle.mvn_sum_logp_grad(X0, mu[None], cov[None]), A_ref)
            X_ref = oracle.svgd_run(X0, iters, mu[None], cov[None], opt_kind=oracle.OPT_ADAM, lr=0.1, scale_method=oracle.SCALE_HESSIAN)
            e_a = np.max(np.abs(Amat - A_ref)) / np.max(np.abs(A_ref))
            e_phi = np.max(np.abs(phi - phi_ref)) / np.max(np.abs(phi_ref))
            e_x = np.sqrt(np.mean((X - X_ref) ** 2)) / np.sqrt(np.mean(X_ref ** 2))
            ok = e_a < 1e-12 and e_phi < 1e-11 and e_x < 1e-9
            print("world=%d f64 Hessian scale n=%d d=%d: A err %.2e, phi err %.2e, trajectory rms err %.2e -> %s" % (world, n, d, e_a, e_phi, e_x, "OK" if ok else "FAIL"), flush=True)
            failures += 0 if ok else 1
    f = torch.tensor([failures], device="cuda")
    dist.all_reduce(f)
    dist.destroy_process_group()
    sys.exit(int(f.item()))


if __name__ == "__main__":
    main()
