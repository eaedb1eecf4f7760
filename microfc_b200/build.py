"""In-tree build of libmfc_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m microfc_b200.build [--force]

The shared library is written next to this file so that it travels with the source snapshot
to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libmfc_b200.so")
# tuning builds (tools/tune_variants.py): extra -D flags and an alternative output name
EXTRA_DEFS = os.environ.get("MFC_B200_DEFS", "").split()
if os.environ.get("MFC_B200_LIBNAME"):
    LIB = os.path.join(HERE, os.environ["MFC_B200_LIBNAME"])
    OBJ = os.path.join(CSRC, "build_" + os.environ["MFC_B200_LIBNAME"].replace(".", "_"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
UNITS = [
    # (source, extra flags)
    ("kernels_fast.cu", []),
    ("kernels_strict.cu", ["-fmad=false"]),
    ("kernels_fast_nf3.cu", []),
    ("kernels_strict_nf3.cu", ["-fmad=false"]),
    ("mfc_api.cu", ["-Xcompiler", "-fvisibility=default"]),
    ("weno_coefficients.cpp", ["-Xcompiler", "-ffp-contract=off"]),
    ("patches.cu", ["-fmad=false"]),
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "mfc_b200.h"))
    headers.append(os.path.abspath(__file__))
    jobs, objs = [], []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + EXTRA_DEFS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    if jobs or force or _stale(LIB, objs):
        run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
