"""Builds libqibo_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m qibo_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the working tree.  nvcc cross-compiles without a GPU.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libqibo_b200.so")
# experiment variants: QB_BUILD_DEFS="-DQB_COMPUTE_THREADS=512" QB_BUILD_SUFFIX=_ct512 python -m qibo_b200.build --force
SOURCES = [os.path.join(CSRC, "qb_api.cu")]


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "qibo_b200.h"))
    return deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=True):
    global LIB
    suffix = os.environ.get("QB_BUILD_SUFFIX", "")
    if suffix:
        LIB = LIB.replace(".so", suffix + ".so")
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [
        nvcc, "-O3", "-std=c++17", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
        "-Xptxas", "-v" if os.environ.get("QB_PTXAS_V") else "-O3",
        "--expt-relaxed-constexpr",
        "-shared", "-o", LIB,
    ] + os.environ.get("QB_BUILD_DEFS", "").split() + SOURCES
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print("built", LIB)
