"""Build libfennec_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

    python -m fennec_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false: nothing is contracted implicitly, so the
bit-exact FP64 paths keep the reference's unfused multiply/add order; FMAs are written explicitly
(fmaf / __ffma2_rn) where they are wanted.  -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libfennec_b200.so")
SOURCES = ["api.cu", "ssim.cu", "box.cu", "effects.cu", "resize.cu", "ycbcr.cu", "analyze.cu", "orient.cu", "palette.cu", "pixfmt.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("FB_EXTRA_NVCC", "").split()
FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "fennec_b200.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}")

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(SO, objs):
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
