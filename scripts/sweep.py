"""BASELINE configs[4]: scaling sweep N = 1K .. 1M, d = 2 .. 512 at 1 / 2 / 4 / 8 GPUs (device-timed steps, X resident, median
bandwidth every step, Adam, the config-3 MVN recipe).  TC32 for d <= 256 (and FP64 beside it where that is cheap), FP64 (DMMA)
for d = 512.  Cells predicted to take longer than `--budget` seconds are skipped.  One JSON line per cell:

    python scripts/sweep.py > profiles/r02/sweep_1gpu.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep.py --ns 65536,262144,1048576 > ...

`--cpu` adds, once per dimension, the two OpenMP ports of the reference path (reference-shaped and blocked Gram-form) on a bounded
particle sample, with the core count (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
from svgdcpp_b200 import synth, _capi  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def cell(lib, dist, torch, n, d, precision, steps, warmup):
    x0, means, covs = synth.mvn_problem(n, d)
    X = np.ascontiguousarray(x0.T)
    dp = C.POINTER(C.c_double)
    ctx = C.c_void_p()

    def check(rc):
        if rc != 0:
            raise RuntimeError(lib.svgdb_last_error(ctx).decode())

    check(lib.svgdb_create(C.byref(ctx), LOCAL, n, d, precision))
    try:
        if WORLD > 1:
            uid = np.zeros(128, dtype=np.uint8)
            if RANK == 0:
                assert lib.svgdb_nccl_unique_id(uid.ctypes.data_as(C.c_void_p), 128) == 0
            t = torch.from_numpy(uid).cuda()
            dist.broadcast(t, 0)
            uid = t.cpu().numpy()
            check(lib.svgdb_comm_init(ctx, WORLD, RANK, uid.ctypes.data_as(C.c_void_p), 128))
        m_, c_ = np.ascontiguousarray(means), np.ascontiguousarray(covs)
        check(lib.svgdb_set_model_mvn(ctx, m_.ctypes.data_as(dp), c_.ctypes.data_as(dp)))
        check(lib.svgdb_set_kernel_rbf(ctx, _capi.SCALE_MEDIAN, 0.0))
        check(lib.svgdb_set_optimizer(ctx, _capi.OPT_ADAM, 0.1, 0.9, 0.999, 1e-8))
        check(lib.svgdb_set_particles(ctx, X.ctypes.data_as(dp)))
        check(lib.svgdb_initialize(ctx))
        check(lib.svgdb_step(ctx, warmup))
        check(lib.svgdb_sync(ctx))
        st0 = _capi.Stats()
        check(lib.svgdb_get_stats(ctx, C.byref(st0)))
        if WORLD > 1:
            dist.barrier()
        ms = C.c_float()
        check(lib.svgdb_time_steps(ctx, steps, C.byref(ms)))
        st = _capi.Stats()
        check(lib.svgdb_get_stats(ctx, C.byref(st)))
        per = ms.value / steps
        if WORLD > 1:
            tt = torch.tensor([per], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            per = float(tt.item())
    finally:
        lib.svgdb_destroy(ctx)
    return {"n": n, "d": d, "gpus": WORLD, "precision": "tc32" if precision == 1 else "f64", "ms_per_step": per,
            "pairs_per_s": float(n) * n / (per * 1e-3), "algorithmic_tflops": (6 * d + 2) * float(n) * n / (per * 1e-3) * 1e-12, "steps": steps,
            "median_passes_per_step": (st.median_passes - st0.median_passes) / steps}


def cpu_cells(d):
    import oracle_binding as oracle

    threads = max(1, len(os.sched_getaffinity(0)))
    out = []
    for shape, n in (("refshape", 1024 if d > 64 else 2048), ("blocked", 4096)):
        x0, means, covs = synth.mvn_problem(n, d)
        X = np.ascontiguousarray(x0.T)
        secs, _ = oracle.timed_iterations(X, 1, means, covs, shape=shape, threads=threads, opt_kind=oracle.OPT_ADAM, lr=0.1)
        out.append({"cpu_port": "reference-shaped" if shape == "refshape" else "blocked Gram-form", "d": d, "sample_particles": n, "cores": threads,
                    "pairs_per_s": n * n / secs})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=20.0, help="skip cells predicted to take longer than this many seconds")
    ap.add_argument("--ns", default="1024,4096,16384,65536,262144,1048576")
    ap.add_argument("--ds", default="2,8,32,64,128,256,512")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--no-f64", action="store_true", help="FP64 only where the tensor-core path does not reach (d = 512)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    lib = _capi.load()
    # crude time model (seconds per step on one GPU) from the measured shapes: TC32 2.6 ms at N = 65536, d = 64 and 500 ms at
    # N = 262144, d = 256 (~ (d / 64)^1.3 per pair); FP64 80 ms at N = 65536, d = 64
    for d in [int(x) for x in args.ds.split(",")]:
        if args.cpu and RANK == 0:
            for c in cpu_cells(d):
                print(json.dumps(c), flush=True)
        for n in [int(x) for x in args.ns.split(",")]:
            precisions = [1] if d <= 256 else []
            if d > 256 or not args.no_f64:
                precisions.append(0)
            for precision in precisions:
                pairs = (n / 65536.0) ** 2
                est = (2.6e-3 * max(d, 16) / 64.0 * max(1.0, d / 64.0) ** 0.3 if precision == 1 else 80e-3 * max(d, 16) / 64.0 * max(1.0, d / 128.0)) * pairs / WORLD
                steps, warmup = 5, 8
                if est * (steps + warmup) * 1.3 > args.budget:
                    if RANK == 0:
                        print(json.dumps({"n": n, "d": d, "gpus": WORLD, "precision": "tc32" if precision == 1 else "f64", "skipped": "estimated %.2f s per step" % est}), flush=True)
                    continue
                t0 = time.time()
                try:
                    out = cell(lib, dist, torch, n, d, precision, steps, warmup)
                except Exception as e:  # report and go on
                    out = {"n": n, "d": d, "gpus": WORLD, "precision": "tc32" if precision == 1 else "f64", "error": str(e)}
                out["wall_s"] = time.time() - t0
                if RANK == 0:
                    print(json.dumps(out), flush=True)
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
