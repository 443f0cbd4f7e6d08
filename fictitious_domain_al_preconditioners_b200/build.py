"""In-tree build of the CUDA library (sm_100a) and the host setup helper.

``python -m fictitious_domain_al_preconditioners_b200.build`` — nvcc cross-compiles
without a GPU.  The result ``csrc/libfdal.so`` travels with the repo snapshot to
the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "libfdal.so")
SOURCES = ["fdal.cu", "setup.cu", "bsr_build.cu"]
HEADERS = ["kernels.cuh", "host_finalize.h", "bsr_build.h", os.path.join("..", "..", "include", "fdal.h")]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False, with_nccl: bool | None = None) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    if with_nccl is None:
        with_nccl = os.path.exists(os.path.join(CSRC, "comm.cu"))
    if with_nccl:
        srcs.append(os.path.join(CSRC, "comm.cu"))
        deps.append(os.path.join(CSRC, "comm.cu"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-fopenmp",
        "-ccbin", "/usr/bin/g++",
        "-shared", "-o", LIB,
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-lgomp", "-ldl"]
    cmd += srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libfdal.so")
    if verbose:
        print(r.stdout + r.stderr)
    from .amg_setup import build_host_lib

    build_host_lib(force=force)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
