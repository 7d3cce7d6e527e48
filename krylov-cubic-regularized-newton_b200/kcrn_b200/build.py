"""In-tree build of libkcrn.so (hand-written CUDA for sm_100a, C ABI in include/kcrn.h).

``python -m kcrn_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU;
the shared object lands next to this file so that it travels with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(PKG))
CSRC = os.path.join(os.path.dirname(PKG), "csrc")
OBJ = os.path.join(os.path.dirname(PKG), "build")
LIB = os.path.join(PKG, "libkcrn.so")
UNITS = ["kcrn_api", "kcrn_transpose", "kcrn_sweep", "kcrn_lanczos", "kcrn_cubic", "kcrn_config4", "kcrn_nccl", "kcrn_run", "kcrn_batch", "kcrn_svmlight", "kcrn_persist"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libkcrn.so cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "kcrn.h"))
    nvcc = _nvcc()
    jobs = []
    for u in UNITS:
        src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(OBJ, u + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + p.stdout + p.stderr)
        return p.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for log in ex.map(run, jobs):
            if verbose and log:
                print(log, file=sys.stderr)
    objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
    if force or jobs or _stale(LIB, objs):
        run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
