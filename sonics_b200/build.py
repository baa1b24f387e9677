"""Compile the CUDA library in-tree: sonics_b200/libsonics_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsonics_b200.so")
SOURCES = ["sonics_b200.cu"]
DEPS = ["sonics_b200.cu", "sim_kernel.cuh", "sim_lockstep.cuh", "finish.cuh", "stats_kernel.cuh", "abi_stats.inl", "host_io.inl", "samplers.cuh", "rng.cuh", "lnl.cuh",
        os.path.join("..", "..", "include", "sonics_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wno-deprecated-declarations",
    "-shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
