"""Builds libsvgd_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

`python -m svgdcpp_b200.build` or `svgdcpp_b200.build.build()`.  The library is written to
svgdcpp_b200/lib/ so it travels with the source tree; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsvgd_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> list[str]:
    # the image exports CC/CXX pointing at a wrapper without libgomp; use the system g++
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".hpp")))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(HERE, "..", "include", "svgd_b200.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    defines = ["-DSVGDB_WITH_TC32"] if os.path.exists(os.path.join(CSRC, "kernels_tc32.cuh")) else []
    cmd = [_nvcc(), *NVCC_FLAGS, *_host_cxx(), *defines, "-o", LIB_PATH,
           os.path.join(CSRC, "svgd_b200_api.cu"), "-ldl", "-lcuda"]
    if trace:
        cmd.insert(1, "-DSVGDB_TC_TRACE_BUILD")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    # --trace: compile the SVGDB_TC_TRACE timeline probes into the pair kernel (development aid, scripts/tc_trace.py)
    print(build(force="--force" in sys.argv or "--trace" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
