#!/usr/bin/env python
"""bench.py — headline benchmark of the SVGD inner loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f64|tc32]

Workload: BASELINE.json configs[2] — 64-D MVN with dense covariance, N = 65,536 particles, median-
heuristic bandwidth recomputed every iteration, Adam(0.1, 0.9, 0.999); synthetic inputs from
svgdcpp_b200.synth (splitmix64 + Box-Muller).  One "step" = one SVGD::Step
(reference SVGD.hpp:373-400) over all N^2 ordered particle pairs.  metric = N^2 * steps / seconds.

N > 1 (torchrun, one rank per GPU): the same particle set, rows sharded over the ranks, NCCL
all-gather of X and V per step (strong scaling).  Timing: CUDA events on the stream the kernels run on,
barrier + synchronize on both sides, max over ranks.

`--impl reference` times the reference's own algorithm on the host cores.  SVGDCpp cannot be built
here (it needs Eigen + CppAD; neither is installed and there is no network), so that arm runs the
reference-shaped OpenMP port in oracle/ (cpu_baseline.kind == "port") on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_PARTICLES = 65536
DIM = 64
METRIC = "particle-pair interactions/sec (N^2*iters/s) at N=65536,d=64"
UNIT = "pairs/s"
CPU_SAMPLE_N = 2048
CPU_SAMPLE_ITERS = 3


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p.get("bf16_tflops"), "bf16_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(shape="refshape", iters=CPU_SAMPLE_ITERS, n=CPU_SAMPLE_N):
    """The reference-shaped OpenMP port of the oracle on a bounded sample of the workload."""
    import oracle_binding as oracle
    from svgdcpp_b200 import synth

    x0, means, covs = synth.mvn_problem(N_PARTICLES, DIM)
    X = np.ascontiguousarray(x0.T[:n])
    threads = oracle.max_threads()
    secs, _ = oracle.timed_iterations(X, iters, means, covs, shape=shape, threads=threads, opt_kind=oracle.OPT_ADAM, lr=0.1)
    return {"value": n * n * iters / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d particles of the N=%d, d=%d set, %d iterations, %s OpenMP port of SVGD.hpp:373-454 "
                      "(Eigen/CppAD reference not buildable here)" % (n, N_PARTICLES, DIM, iters,
                                                                     "reference-shaped (K, grad K materialised, indexer GEMM)" if shape == "refshape" else "blocked Gram-form"),
            "seconds": secs}


def run_reference(args, rank):
    if rank != 0:
        return
    t_all = []
    for _ in range(args.warmup):
        cpu_baseline(iters=1, n=1024)
    for _ in range(args.steps):
        t_all.append(cpu_baseline(iters=1))
    secs = sum(b["seconds"] for b in t_all)
    value = CPU_SAMPLE_N * CPU_SAMPLE_N * len(t_all) / secs
    base = t_all[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(t_all), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "64-D MVN dense covariance, N=65536, median bandwidth, Adam (BASELINE configs[2])",
                   "n_particles": N_PARTICLES, "dim": DIM, "sample_particles": CPU_SAMPLE_N},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": base["cores"], "kind": "port",
                         "sample": base["sample"].replace("%d iterations" % 1, "1 iteration per step")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SVGDB_BENCH_PRECISION", "auto"), choices=["auto", "f64", "tc32"])
    ap.add_argument("--particles", type=int, default=N_PARTICLES)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import svgdcpp_b200 as sv
    from svgdcpp_b200 import _capi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SVGD path has no CPU fallback")
    lib = _capi.load()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, d = args.particles, args.dim
    x0, means, covs = synth.mvn_problem(n, d)
    nbytes = n * d * 8

    # pinned host staging buffer holding the particle matrix in the reference layout (d x n column-major)
    hp = C.c_void_p()
    assert lib.svgdb_host_alloc(C.byref(hp), nbytes) == 0
    host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(n, d))
    host[...] = x0.T

    precision = _capi.PRECISION_F64
    if args.precision == "tc32" or (args.precision == "auto" and os.path.exists(os.path.join(ROOT, "svgdcpp_b200", "csrc", "kernels_tc32.cuh"))):
        precision = _capi.PRECISION_TC32
    ctx = C.c_void_p()

    def check(rc):
        if rc != 0:
            raise RuntimeError("svgd_b200: %s" % lib.svgdb_last_error(ctx).decode())

    check(lib.svgdb_create(C.byref(ctx), local_rank, n, d, precision))
    stream = torch.cuda.current_stream()
    check(lib.svgdb_set_stream(ctx, C.c_void_p(stream.cuda_stream)))
    if world > 1:
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            assert lib.svgdb_nccl_unique_id(uid.ctypes.data_as(C.c_void_p), 128) == 0
        t = torch.from_numpy(uid).cuda()
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
        check(lib.svgdb_comm_init(ctx, world, rank, uid.ctypes.data_as(C.c_void_p), 128))
    dp = C.POINTER(C.c_double)
    m_, c_ = np.ascontiguousarray(means), np.ascontiguousarray(covs)
    check(lib.svgdb_set_model_mvn(ctx, m_.ctypes.data_as(dp), c_.ctypes.data_as(dp)))
    check(lib.svgdb_set_kernel_rbf(ctx, _capi.SCALE_MEDIAN, 0.0))
    check(lib.svgdb_set_optimizer(ctx, _capi.OPT_ADAM, 0.1, 0.9, 0.999, 1e-8))
    check(lib.svgdb_set_particles(ctx, host.ctypes.data_as(dp)))
    check(lib.svgdb_initialize(ctx))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stats():
        st = _capi.Stats()
        check(lib.svgdb_get_stats(ctx, C.byref(st)))
        return st

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
    ms = e0.elapsed_time(e1)
    st = stats()
    launches = int(st.kernel_launches)
    median_passes = int(st.median_passes)
    clocks = sampler.summary() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- per-kernel leg for the roofline: CUDA events around the pair-interaction kernel ----------
    check(lib.svgdb_reset_stats(ctx))
    check(lib.svgdb_set_profiling(ctx, 1))
    check(lib.svgdb_step(ctx, max(2, min(args.steps, 5))))
    barrier()
    sp = stats()
    check(lib.svgdb_set_profiling(ctx, 0))
    phi_phase_ms = sp.ms_phi / max(1, sp.phi_launches)           # operand preparation + kernel + optimizer kernel
    phi_ms = sp.ms_phi_kernel / max(1, sp.phi_launches)          # the pair-interaction kernel alone (events around its launch)
    if not phi_ms > 0.0:
        phi_ms = phi_phase_ms
    prof_iters = max(1, int(sp.iterations))
    phase_ms = {"median": sp.ms_median / prof_iters, "grad": sp.ms_grad / prof_iters, "phi": sp.ms_phi / prof_iters,
                "comm_and_misc": sp.ms_comm / prof_iters}

    # ---- end-to-end leg: host buffers, H2D + step + D2H inside the timed region ------------------
    # Every rank moves ITS rows of the particle matrix (svgdb_set_particles_rows / _get_particles_rows: with one rank these
    # are the whole matrix); the other rows arrive over NVLink.  Bytes per step are summed over the ranks.
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
        # H2D of this step's particles (this rank's rows), one SVGD step, D2H of the result; synchronous.  The D2H of rows
        # that are already updated overlaps the rest of the pair kernel (four row chunks), the bytes moved are the same.
        check(lib.svgdb_step_host(ctx, mine_p, mine_p, 1))
    f1.record(stream)
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t_wall) if world == 1 else 0.0)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    finite = bool(np.all(np.isfinite(mine)))

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
        if os.path.exists(tpath) and world == 1 and n == N_PARTICLES and d == DIM:
            with open(tpath) as f:
                traffic = json.load(f).get("tc32_phi" if precision == _capi.PRECISION_TC32 else "f64_phi")
        roof = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["bf16_sustained"], "traffic": traffic,
                "kernel": "phi2_tc32_kernel (pair interaction)" if precision == _capi.PRECISION_TC32 else "phi_f64_kernel (pair interaction + optimizer epilogue)",
                "kernel_ms": phi_ms, "phase_ms_with_operand_prep_and_optimizer": phi_phase_ms,
                "algorithmic_flops_per_launch": phi_flops, "peak_source": peaks["source"] + ", dense bf16 sustained",
                "whole_step_algorithmic_tflops": step_tf / max(world, 1), "phase_ms_per_step": phase_ms}
        if precision == _capi.PRECISION_F64:
            dm = C.c_double(0.0)
            if lib.svgdb_probe_peak(local_rank, 0, C.byref(dm)) == 0 and dm.value > 0:
                roof["issued_kind"] = "fp64 DMMA (mma.sync.m8n8k4.f64)"
                roof["issued_kind_peak"] = dm.value
                roof["issued_kind_peak_source"] = "measured here by svgdb_probe_peak (register-resident DMMA loop)"
                roof["frac_of_issued_kind"] = achieved_tf / dm.value
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if precision == _capi.PRECISION_F64 else "f32",
            "data": "synthetic",
            "config": {"workload": "64-D MVN dense covariance, N=%d, median bandwidth, Adam (BASELINE configs[2])" % n,
                       "n_particles": n, "dim": d, "parallelism": "rows sharded over %d GPU(s), NCCL all-gather of X and V" % world,
                       "l2": "working set (X, V, X_next, optimizer state) = %d MB > 126 MB L2; compute-bound, no flush" % (5 * nbytes // 2 ** 20),
                       "median_passes_per_step": median_passes / max(1, args.steps), "finite": finite,
                       "precision_mode": "F64: DMMA fp64 end to end" if precision == _capi.PRECISION_F64 else
                       "TC32: tcgen05 kind::f16 MMAs on split fp16 (pair kernel) / bf16 (median) particles and scaled-fp16 kernel values, "
                       "fp32 accumulation in TMEM, fp32 ex2, FP64 optimizer state (error bound in DESIGN.md)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "ms_per_step": e2e_ms / args.steps,
                    "call": "svgdb_step_host(rows_in, rows_out, 1) per rank == svgdb_set_particles_rows(host) + svgdb_step(1) + svgdb_get_particles_rows(host) == SVGD::Step() of the facade on a host matrix; bytes summed over ranks"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline().items() if k != "seconds"}
            opt = cpu_baseline(shape="blocked", iters=1, n=8192)
            line["cpu_baseline"]["optimised_port_value"] = opt["value"]
            line["cpu_baseline"]["optimised_port_sample"] = opt["sample"]
        print(json.dumps(line), flush=True)

    lib.svgdb_destroy(ctx)
    lib.svgdb_host_free(hp)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
