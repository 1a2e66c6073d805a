"""Builds libp3cudaskin.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

-fmad=false is part of the numerical contract (csrc/p3cs_math.cuh): the kernels evaluate the
reference's scalar formulas without FMA contraction so results match the CPU path bit for bit.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libp3cudaskin.so")
SOURCES = ["p3cs_api.cu", "p3cs_mesh.cu", "p3cs_interop.cu", "p3cs_kernels.cu", "p3cs_pose.cu", "p3cs_format.cu", "p3cs_crowd.cu", "p3cs_strip.cu",
           "p3cs_run_inst_a.cu", "p3cs_run_inst_b.cu",
           "p3cs_strip_inst_a.cu", "p3cs_strip_inst_b.cu", "p3cs_strip_inst_c.cu", "p3cs_strip_inst_d.cu", "p3cs_strip_inst_e.cu"]
HEADERS = ["p3cs_internal.h", "p3cs_kernels.cuh", "p3cs_math.cuh", "p3cs_device.cuh", "p3cs_strip.cuh", "p3cs_run.cuh", "p3cs_runplan.h",
           os.path.join("..", "..", "include", "p3cudaskin.h")]
OBJ_DIR = os.path.join(HERE, "_obj")   # git-ignored and gpurun-ignored: only the .so travels

COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xfatbin", "-compress-all",
] + [f"-D{d}" for d in os.environ.get("P3CS_DEFINES", "").split() if d]   # experiments: P3CS_DEFINES="P3CS_RUN_CTAS_PER_SM=3"
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libp3cudaskin.so")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every translation unit in parallel (the strip kernel's modes are spread over several), then links."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = [r for _, r in results if r.returncode != 0]
    if failed:
        for r in failed:
            sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libp3cudaskin.so")
    if verbose:
        for _, r in results:
            sys.stderr.write(r.stderr)
    res = subprocess.run([nvcc] + LINK_FLAGS + ["-o", LIB] + [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libp3cudaskin.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
