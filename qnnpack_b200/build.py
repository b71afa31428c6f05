"""Builds libqnnpack.so (the C-ABI product library) in-tree with nvcc for sm_100a.

The shared object lands in qnnpack_b200/lib/ (git-ignored, but shipped to the GPU box by gpurun).
nvcc cross-compiles without a GPU, so this also runs in the CPU-only authoring container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqnnpack.so")

SOURCES = ["qnnpack_api.cu", "q8_igemm_sm100.cu", "q8_dwconv_sm100.cu", "q8_dwconv_stream_sm100.cu",
           "q8_dwconv_umma_sm100.cu", "q8_peak_sm100.cu", "q8_eltwise_sm100.cu"]
HEADERS = ["q8_igemm_sm100.cuh", "q8_dwconv_sm100.cuh", "q8_eltwise_sm100.cuh", "sm100_ptx.cuh", "requant_math.h", "requant_dev.cuh",
           os.path.join("..", "..", "include", "qnnpack.h"), os.path.join("..", "..", "include", "qnnpack_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--shared", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("QNNP_BUILD_DEFINES", "").split()  # e.g. "-DQ8_MBAR_HINT_NS=2000" for experiments
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libqnnpack.so")
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
