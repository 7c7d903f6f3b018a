"""BASELINE configs[4]: scaling sweep N = 1K .. 1M, d = 2 .. 512 on one GPU (device-timed steps, X resident).
TC32 for d <= 64, FP64 (DMMA) for every d.  Cells that would take longer than `--budget` seconds are skipped.
Writes one JSON line per cell; `python scripts/sweep.py > profiles/r01/sweep.jsonl`."""
import argparse
import ctypes as C
import json
import sys
import time

sys.path.insert(0, ".")
import svgdcpp_b200 as sv
from svgdcpp_b200 import synth, _capi


def cell(n, d, precision, steps, warmup):
    x0, means, covs = synth.mvn_problem(n, d)
    model = sv.MultivariateNormal(means[0], covs[0])
    s = sv.SVGD(d, 1, x0, sv.GaussianRBFKernel(x0, sv.ScaleMethod.Median, model), model, sv.Adam(d, n, 0.1, 0.9, 0.999), precision=precision)
    s.Initialize(); s._upload()
    lib = _capi.load()
    rc = lib.svgdb_step(s._ctx, warmup)
    if rc != 0:
        raise RuntimeError(lib.svgdb_last_error(s._ctx).decode())
    before = s.Stats()
    ms = C.c_float()
    rc = lib.svgdb_time_steps(s._ctx, steps, C.byref(ms))
    if rc != 0:
        raise RuntimeError(lib.svgdb_last_error(s._ctx).decode())
    st = s.Stats()
    s.close()
    per = ms.value / steps
    return {"n": n, "d": d, "precision": "tc32" if precision == 1 else "f64", "ms_per_step": per, "pairs_per_s": float(n) * n / (per * 1e-3),
            "algorithmic_tflops": (6 * d + 2) * float(n) * n / (per * 1e-3) * 1e-12, "steps": steps,
            "median_passes_per_step": (st["median_passes"] - before["median_passes"]) / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=20.0, help="skip cells predicted to take longer than this many seconds")
    ap.add_argument("--ns", default="1024,4096,16384,65536,262144,1048576")
    ap.add_argument("--ds", default="2,8,32,64,128,256,512")
    args = ap.parse_args()
    # crude time model (seconds per step) from the headline shape: TC32 3 ms, F64 97 ms at N = 65536, d = 64
    for d in [int(x) for x in args.ds.split(",")]:
        for n in [int(x) for x in args.ns.split(",")]:
            for precision in ((1, 0) if d <= 64 else (0,)):
                scale = (n / 65536.0) ** 2 * max(d, 16) / 64.0
                est = (3e-3 if precision == 1 else 97e-3) * scale
                steps, warmup = 5, 3
                if est * (steps + warmup) * 1.5 > args.budget:
                    print(json.dumps({"n": n, "d": d, "precision": "tc32" if precision == 1 else "f64", "skipped": "estimated %.1f s per step" % est}), flush=True)
                    continue
                t0 = time.time()
                try:
                    out = cell(n, d, precision, steps, warmup)
                except Exception as e:  # report and go on
                    out = {"n": n, "d": d, "precision": "tc32" if precision == 1 else "f64", "error": str(e)}
                out["wall_s"] = time.time() - t0
                print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
