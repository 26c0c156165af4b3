"""In-tree build of the CUDA library (sm_100a only).  `python sci-solver_fem_b200/build.py [-f]`.

Produces sci-solver_fem_b200/libfemsolver_b200.so next to this file so that it travels to the GPU
box with the repository snapshot.  Per-file flags matter: assembly.cu and hierarchy.cu are compiled
with -fmad=false (see the notes at the top of those files).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfemsolver_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
METIS = os.environ.get("METIS_LIB", "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a")  # METIS 5 (aggregatorType_ = 1)
EXTRA = os.environ.get("NVCC_EXTRA", "").split()
COMMON = EXTRA + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
          "-Xcudafe", "--diag_suppress=177", "-Xcompiler", "-Wno-deprecated-declarations"]
SOURCES = {
    "prims.cu": [],
    "pattern.cu": [],
    "assembly.cu": ["-fmad=false"],
    "quadrature.cpp": [],
    "metis_agg.cpp": [],
    "aggregation.cu": [],
    "hierarchy.cu": ["-fmad=false"],
    "cycle.cu": [],
    "dense_tail.cu": [],
    "solver.cu": [],
    "dist.cu": [],
    "capi.cu": [],
}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "femsolver_b200.h"))
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        op = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + ["-x", "cu", "-c", sp, "-o", op]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and r.stderr.strip():
            print(r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(LIB):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + [METIS, "-lcudart", "-lcublas"])
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
