#!/usr/bin/env python
"""bench.py — headline benchmark of the SVGD inner loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f64|tc32] [--workload c3|c4]

Workloads (BASELINE.json configs):
  c3 (default, the configuration the metric is quoted on): 64-D MVN with dense covariance, N = 65,536 particles, median-
     heuristic bandwidth recomputed every iteration, Adam(0.1, 0.9, 0.999).
  c4: 256-D 16-component sum of Gaussians, N = 262,144 particles, AdaGrad(0.1), particles sharded over the ranks.
Synthetic inputs from svgdcpp_b200.synth (splitmix64 + Box-Muller).  One "step" = one SVGD::Step (reference
SVGD.hpp:373-400) over all N^2 ordered particle pairs.  metric = N^2 * steps / seconds.

N > 1 (torchrun, one rank per GPU): the same particle set, rows sharded over the ranks, NCCL all-gather of X and V per step
(strong scaling).  Timing: CUDA events on the stream the kernels run on, barrier + synchronize on both sides, max over ranks.

`--impl reference` times the reference's own algorithm on the host cores.  SVGDCpp cannot be built here (it needs Eigen +
CppAD; neither is installed and there is no network), so that arm runs the reference-shaped OpenMP port in oracle/
(cpu_baseline.kind == "port") on a bounded sample, with every core the process may use.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

UNIT = "pairs/s"
CPU_SAMPLE_N = 2048
CPU_SAMPLE_ITERS = 3

WORKLOADS = {
    "c3": {"n": 65536, "d": 64, "components": 1, "opt": "adam",
           "name": "64-D MVN dense covariance, N=%d, median bandwidth, Adam (BASELINE configs[2])",
           "metric": "particle-pair interactions/sec (N^2*iters/s) at N=65536,d=64"},
    "c4": {"n": 262144, "d": 256, "components": 16, "opt": "adagrad",
           "name": "256-D 16-component sum of Gaussians, N=%d, median bandwidth, AdaGrad, particles sharded over the GPUs (BASELINE configs[3])",
           "metric": "particle-pair interactions/sec (N^2*iters/s) at N=262144,d=256,C=16"},
}


def host_threads() -> int:
    """Every core this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 into its children; the CPU arm is
    meant to use all host threads, so the OpenMP thread count is set explicitly from the affinity mask."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p.get("bf16_tflops"), "bf16_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks, power and throttle reasons through NVML while the timed region runs (one sample every few ms; the
    timed region of the default run lasts ~50 ms, far less than one nvidia-smi process start)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            self.stop_flag.wait(0.004)

    def summary(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = set()
        for s in self.samples:
            for k, bit in names.items():
                if s[3] & bit:
                    reasons.add(k)
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": float(max(s[1] for s in self.samples)),
                "power_w_max": float(max(s[2] for s in self.samples)), "reasons": sorted(reasons), "samples": len(self.samples)}


def problem(workload, n=None, d=None):
    from svgdcpp_b200 import synth

    w = WORKLOADS[workload]
    n, d = n or w["n"], d or w["d"]
    if w["components"] == 1:
        return synth.mvn_problem(n, d)
    return synth.gmm_problem(n, d, w["components"])


def cpu_baseline(workload="c3", shape="refshape", iters=CPU_SAMPLE_ITERS, n=CPU_SAMPLE_N):
    """The reference-shaped OpenMP port of the oracle on a bounded sample of the workload."""
    import oracle_binding as oracle

    w = WORKLOADS[workload]
    x0, means, covs = problem(workload, n=max(n, 4096) if w["components"] > 1 else w["n"])
    X = np.ascontiguousarray(x0.T[:n])
    threads = host_threads()
    kind = oracle.OPT_ADAM if w["opt"] == "adam" else oracle.OPT_ADAGRAD
    secs, _ = oracle.timed_iterations(X, iters, means, covs, shape=shape, threads=threads, opt_kind=kind, lr=0.1)
    return {"value": n * n * iters / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d particles of the %s set (d=%d), %d iterations, %s OpenMP port of SVGD.hpp:373-454 "
                      "(Eigen/CppAD reference not buildable here)" % (n, workload, w["d"], iters,
                                                                     "reference-shaped (K, grad K materialised, indexer GEMM)" if shape == "refshape" else "blocked Gram-form"),
            "seconds": secs}


def run_reference(args, rank):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sample_n = CPU_SAMPLE_N if w["d"] <= 64 else 1024   # grad K is n^2 d doubles: 2 GiB at n = 1024, d = 256
    t_all = []
    for _ in range(args.warmup):
        cpu_baseline(args.workload, iters=1, n=min(1024, sample_n))
    for _ in range(args.steps):
        t_all.append(cpu_baseline(args.workload, iters=1, n=sample_n))
    secs = sum(b["seconds"] for b in t_all)
    value = sample_n * sample_n * len(t_all) / secs
    base = t_all[0]
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(t_all), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"] % w["n"], "n_particles": w["n"], "dim": w["d"], "sample_particles": sample_n,
                   "note": "the reference algorithm needs n^2 (d + 1) doubles (2 TiB at N=65536, d=64): each step runs on a %d-particle sample of the same set; pairs/s is the common unit" % sample_n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": base["cores"], "kind": "port",
                         "sample": base["sample"].replace("%d iterations" % 1, "1 iteration per step")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def sampled_parity(lib, ctx, check, n, d, means, covs, rank, tol):
    """Correctness inside the bench run (all ranks take part in the collectives, rank 0 checks): ComputePhi on the current particles,
    then 8 sampled rows of phi recomputed on the host in FP64 (O(N d) each) from the device's own grad log p, 64 sampled rows of
    grad log p recomputed from the model, and the kernel scale against the median of 2M sampled pair distances."""
    dp = C.POINTER(C.c_double)
    X = np.empty((n, d))
    G = np.empty((n, d))
    phi = np.empty((n, d))
    a = C.c_double(0.0)
    check(lib.svgdb_get_particles(ctx, X.ctypes.data_as(dp)))
    check(lib.svgdb_compute_log_model_grad(ctx, G.ctypes.data_as(dp)))
    check(lib.svgdb_compute_phi(ctx, phi.ctypes.data_as(dp), C.byref(a)))
    if rank != 0:
        return None
    a = a.value
    rng = np.random.default_rng(12345)
    scale = float(np.max(np.abs(phi)))
    worst = 0.0
    for i in rng.integers(0, n, 8):
        diff = X - X[i]
        k = np.exp(-a * np.einsum("ij,ij->i", diff, diff))
        ref = (k @ G + (-2.0 * a * diff * k[:, None]).sum(0)) / n
        worst = max(worst, float(np.max(np.abs(phi[i] - ref)) / scale))
    # grad log p of the sum of unnormalised Gaussians at 64 sampled particles (log-sum-exp form)
    rows = rng.integers(0, n, 64)
    P = np.linalg.inv(covs)
    P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
    diffs = X[rows][None, :, :] - means[:, None, :]                       # C x 64 x d
    Y = np.einsum("cnd,cde->cne", diffs, P)
    q = -0.5 * np.einsum("cnd,cnd->cn", Y, diffs)
    r = np.exp(q - q.max(0))
    r /= r.sum(0)
    g_ref = -np.einsum("cn,cnd->nd", r, Y)
    g_err = float(np.max(np.abs(G[rows] - g_ref)) / max(1e-300, np.max(np.abs(g_ref))))
    # kernel scale: the median of all n^2 distances, estimated from 2M sampled ordered pairs (incl. the n/n^2 share of zeros)
    m = 2_000_000 if d <= 64 else 250_000   # (bounded host memory: m x d doubles twice)
    ii, jj = rng.integers(0, n, m), rng.integers(0, n, m)
    dist = np.sqrt(np.einsum("ij,ij->i", X[ii] - X[jj], X[ii] - X[jj]))
    a_est = np.log(n) / np.median(dist) ** 2
    a_err = abs(a - a_est) / a_est
    return {"phi_sampled_rows_max_err_over_max_phi": worst, "phi_tol": tol, "grad_sampled_rows_rel_err": g_err,
            "scale": a, "scale_rel_diff_vs_sampled_median_estimate": a_err,
            "ok": bool(worst < tol and g_err < 1e-6 and a_err < 2e-2 and np.all(np.isfinite(phi)))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SVGDB_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("SVGDB_BENCH_PRECISION", "auto"), choices=["auto", "f64", "tc32"])
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f64-leg", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from svgdcpp_b200 import _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SVGD path has no CPU fallback")
    lib = _capi.load()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = WORKLOADS[args.workload]
    n, d = args.particles or w["n"], args.dim or w["d"]
    x0, means, covs = problem(args.workload, n, d)
    nbytes = n * d * 8
    is_default_shape = n == w["n"] and d == w["d"]

    # pinned host staging buffer holding the particle matrix in the reference layout (d x n column-major)
    hp = C.c_void_p()
    assert lib.svgdb_host_alloc(C.byref(hp), nbytes) == 0
    host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(n, d))
    host[...] = x0.T

    precision = _capi.PRECISION_TC32 if args.precision in ("tc32", "auto") else _capi.PRECISION_F64
    dp = C.POINTER(C.c_double)
    m_, c_ = np.ascontiguousarray(means), np.ascontiguousarray(covs)
    stream = torch.cuda.current_stream()

    def make_ctx(prec):
        ctx = C.c_void_p()

        def check(rc):
            if rc != 0:
                raise RuntimeError("svgd_b200: %s" % lib.svgdb_last_error(ctx).decode())

        check(lib.svgdb_create(C.byref(ctx), local_rank, n, d, prec))
        check(lib.svgdb_set_stream(ctx, C.c_void_p(stream.cuda_stream)))
        if world > 1:
            uid = np.zeros(128, dtype=np.uint8)
            if rank == 0:
                assert lib.svgdb_nccl_unique_id(uid.ctypes.data_as(C.c_void_p), 128) == 0
            t = torch.from_numpy(uid).cuda()
            dist.broadcast(t, 0)
            uid = t.cpu().numpy()
            check(lib.svgdb_comm_init(ctx, world, rank, uid.ctypes.data_as(C.c_void_p), 128))
        check(lib.svgdb_set_model_mvn_sum(ctx, m_.shape[0], m_.ctypes.data_as(dp), c_.ctypes.data_as(dp)))
        check(lib.svgdb_set_kernel_rbf(ctx, _capi.SCALE_MEDIAN, 0.0))
        if w["opt"] == "adam":
            check(lib.svgdb_set_optimizer(ctx, _capi.OPT_ADAM, 0.1, 0.9, 0.999, 1e-8))
        else:
            check(lib.svgdb_set_optimizer(ctx, _capi.OPT_ADAGRAD, 0.1, 0.0, 0.0, 1e-8))
        check(lib.svgdb_set_particles(ctx, host.ctypes.data_as(dp)))
        check(lib.svgdb_initialize(ctx))
        return ctx, check

    ctx, check = make_ctx(precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stats():
        st = _capi.Stats()
        check(lib.svgdb_get_stats(ctx, C.byref(st)))
        return st

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    # ---- device-resident leg: `value` ------------------------------------------------------------
    check(lib.svgdb_step(ctx, args.warmup))
    barrier()
    check(lib.svgdb_reset_stats(ctx))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    check(lib.svgdb_step(ctx, args.steps))
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    st = stats()
    launches = int(st.kernel_launches)
    median_passes = int(st.median_passes)
    clocks = sampler.summary() if rank == 0 else None

    # ---- per-kernel leg for the roofline: CUDA events around the pair-interaction kernel ----------
    check(lib.svgdb_reset_stats(ctx))
    check(lib.svgdb_set_profiling(ctx, 1))
    check(lib.svgdb_step(ctx, max(2, min(args.steps, 5))))
    barrier()
    sp = stats()
    check(lib.svgdb_set_profiling(ctx, 0))
    phi_phase_ms = sp.ms_phi / max(1, sp.phi_launches)           # operand preparation + kernel (+ optimizer)
    phi_ms = sp.ms_phi_kernel / max(1, sp.phi_launches)          # the pair-interaction kernel alone (events around its launch)
    if not phi_ms > 0.0:
        phi_ms = phi_phase_ms
    prof_iters = max(1, int(sp.iterations))
    phase_ms = {"median": sp.ms_median / prof_iters, "grad": sp.ms_grad / prof_iters, "phi": sp.ms_phi / prof_iters,
                "comm_and_misc": sp.ms_comm / prof_iters,
                "grad_kernel_on_side_stream": sp.ms_grad_kernel / prof_iters, "median_passes": sp.median_passes / prof_iters}

    # ---- correctness of what was just timed (all ranks; rank 0 reports) ---------------------------
    parity = None
    if not args.no_parity:
        parity = sampled_parity(lib, ctx, check, n, d, means, covs, rank, 2e-4 if precision == _capi.PRECISION_TC32 else 1e-9)

    # ---- end-to-end leg: host buffers, H2D + step + D2H inside the timed region ------------------
    # Every rank moves ITS rows of the particle matrix (svgdb_step_host: with one rank these are the whole matrix); the other
    # rows arrive over NVLink.  Bytes per step are summed over the ranks.
    host[...] = x0.T
    check(lib.svgdb_set_particles(ctx, host.ctypes.data_as(dp)))
    check(lib.svgdb_initialize(ctx))
    r0, nr = C.c_int64(0), C.c_int64(0)
    check(lib.svgdb_local_rows(ctx, C.byref(r0), C.byref(nr)))
    mine = host[r0.value:r0.value + nr.value]                       # contiguous view of this rank's rows in the pinned buffer
    mine_p = mine.ctypes.data_as(dp)
    for _ in range(args.warmup):
        check(lib.svgdb_step_host(ctx, mine_p, mine_p, 1))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    f0.record(stream)
    for _ in range(args.steps):
        # H2D of this step's particles (this rank's rows), one SVGD step, D2H of the result; synchronous.
        check(lib.svgdb_step_host(ctx, mine_p, mine_p, 1))
    f1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t_wall) if world == 1 else 0.0))
    finite = bool(np.all(np.isfinite(mine)))
    lib.svgdb_destroy(ctx)

    # ---- equal-precision leg: the FP64 (DMMA) mode on the same workload, a few steps ----------------
    f64_leg = None
    if precision == _capi.PRECISION_TC32 and not args.no_f64_leg and args.workload == "c3":
        host[...] = x0.T
        ctx, check = make_ctx(_capi.PRECISION_F64)
        check(lib.svgdb_step(ctx, 2))
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        check(lib.svgdb_step(ctx, 3))
        g1.record(stream)
        barrier()
        f64_ms = max_over_ranks(g0.elapsed_time(g1)) / 3.0
        lib.svgdb_destroy(ctx)
        f64_leg = {"ms_per_step": f64_ms, "value": float(n) * float(n) / (f64_ms * 1e-3), "unit": UNIT, "steps": 3,
                   "note": "SVGDB_PRECISION_F64: IEEE double end to end (DMMA), the reference's own precision"}

    if rank == 0:
        peaks = load_peaks()
        pairs = float(n) * float(n)
        value = pairs * args.steps / (ms * 1e-3)
        e2e_value = pairs * args.steps / (e2e_ms * 1e-3)
        rows_local = (n + world - 1) // world
        phi_flops = (4 * d + 2) * float(rows_local) * float(n)          # algorithmic flops of ONE launch (this rank's rows)
        achieved_tf = phi_flops / (phi_ms * 1e-3) * 1e-12 if phi_ms > 0 else 0.0
        step_tf = (6 * d + 2) * pairs * args.steps / (ms * 1e-3) * 1e-12  # incl. one distance evaluation for the median
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tpath) and world == 1 and is_default_shape and args.workload == "c3":
            with open(tpath) as f:
                traffic = json.load(f).get("tc32_phi" if precision == _capi.PRECISION_TC32 else "f64_phi")
        tc = precision == _capi.PRECISION_TC32
        # the pair kernel is timed inside a short profiling leg (a few steps): the burst peak is its denominator
        roof = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["bf16_burst"], "frac_burst": achieved_tf / peaks["bf16_burst"],
                "frac_sustained": achieved_tf / peaks["bf16_sustained"], "peak_sustained": peaks["bf16_sustained"], "traffic": traffic,
                "kernel": "phi tensor-core pair-interaction kernel (tcgen05)" if tc else "phi_f64_kernel (pair interaction + optimizer epilogue)",
                "kernel_ms": phi_ms, "phase_ms_with_operand_prep": phi_phase_ms,
                "algorithmic_flops_per_launch": phi_flops, "peak_source": peaks["source"] + ", dense bf16 burst (kernel timed over a few steps); sustained beside it",
                "whole_step_algorithmic_tflops": step_tf / max(world, 1),
                "whole_step_frac_burst": step_tf / max(world, 1) / peaks["bf16_burst"],
                "whole_step_frac_sustained": step_tf / max(world, 1) / peaks["bf16_sustained"],
                "phase_ms_per_step": phase_ms}
        if not tc:
            dm = C.c_double(0.0)
            if lib.svgdb_probe_peak(local_rank, 0, C.byref(dm)) == 0 and dm.value > 0:
                roof["issued_kind"] = "fp64 DMMA (mma.sync.m8n8k4.f64)"
                roof["issued_kind_peak"] = dm.value
                roof["issued_kind_peak_source"] = "measured here by svgdb_probe_peak (register-resident DMMA loop)"
                roof["frac_of_issued_kind"] = achieved_tf / dm.value
        line = {
            "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16x2-split/f32-acc (tcgen05 kind::f16 + an e5m2 correction term; fp16 kernel values; FP64 optimizer state)" if tc else "f64",
            "data": "synthetic",
            "config": {"workload": w["name"] % n,
                       "n_particles": n, "dim": d, "parallelism": "rows sharded over %d GPU(s), NCCL all-gather of the operands per step" % world,
                       "l2": "working set (X, V, X_next, optimizer state) = %d MB > 126 MB L2; compute-bound, no flush" % (5 * nbytes // 2 ** 20),
                       "median_passes_per_step": median_passes / max(1, args.steps), "finite": finite,
                       "precision_mode": "F64: DMMA fp64 end to end" if not tc else
                       "TC32: tcgen05 kind::f16 MMAs on split fp16 (pair kernel; the row particle's second term as an e5m2 kind::f8f6f4 product at d >= 48) / "
                       "fp16 two-product or bf16x3 (median) particles and scaled-fp16 kernel values, "
                       "fp32 accumulation in TMEM, fp32 ex2, FP64 optimizer state (error bound in DESIGN.md)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "ms_per_step": e2e_ms / args.steps,
                    "call": "svgdb_step_host(rows_in, rows_out, 1) per rank == svgdb_set_particles_rows(host) + svgdb_step(1) + svgdb_get_particles_rows(host) == SVGD::Step() of the facade on a host matrix; bytes summed over ranks"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "parity": parity,
        }
        if f64_leg is not None:
            line["f64_mode"] = f64_leg
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.workload).items() if k != "seconds"}
            if args.workload == "c3":
                opt = cpu_baseline(args.workload, shape="blocked", iters=1, n=8192)
                line["cpu_baseline"]["optimised_port_value"] = opt["value"]
                line["cpu_baseline"]["optimised_port_sample"] = opt["sample"]
        print(json.dumps(line), flush=True)

    lib.svgdb_host_free(hp)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
