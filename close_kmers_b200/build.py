"""Build libckm.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.  No torch involved."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libckm.so")
KSER = os.path.join(HERE, "kser_b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["csrc/ckm_api.cu", "host/handlers.cc", "host/seq_parser.cc", "host/http.cc", "host/lookup.cc", "host/kser.cc"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler",
         "-fPIC,-Wall,-Wno-unused-function", "-cudart", "static", "-lz", "-lpthread"]


def _newest_source() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "host"), os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h", ".cc", ".c")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.exists(KSER) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit would still carry the prebuilt library
        raise RuntimeError(f"nvcc not found at {NVCC} and {LIB} is missing")
    extra = os.environ.get("CKM_NVCC_EXTRA", "").split()  # e.g. -DCKM_EXPERIMENTS: the A/B kernels and tuning bits (tools/)
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed")
    # the server binary: main() only, everything else lives in the library next to it
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-o", KSER, os.path.join(HERE, "host", "kser_main.c"), "-L" + HERE, "-lckm",
           "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking kser_b200 failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
