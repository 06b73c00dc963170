"""In-tree build of libbartint_cuda.so (sm_100a) and the CPU-only synthetic-forest generator.

nvcc cross-compiles without a GPU; the .so files stay next to the sources (git-ignored) so they
travel to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbartint_cuda.so")
SYNTH_LIB = os.path.join(HERE, "libbartint_synth.so")

CUDA_SOURCES = ["bartint_abi.cu", "forest_host.cu", "kernels_eval.cu", "kernels_gather.cu", "kernels_bitslice.cu",
                "kernels_misc.cu", "group.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-DBARTINT_TEST_HOOKS",       # host-only hooks of include/bartint_test_hooks.h (the CPU test-suite needs them)
    "-Xcompiler", "-fPIC,-fopenmp,-O3,-pthread", "-shared", "-lgomp", "-ldl",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libbartint_cuda.so cannot be built (there is no CPU fallback)")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "pipeline.cuh"),
                   os.path.join(HERE, "..", "include", "bartint.h"),
                   os.path.join(HERE, "..", "include", "bartint_test_hooks.h")]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, *(extra or []), "-o", LIB, *srcs]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB


def build_synth(force: bool = False) -> str:
    src = os.path.join(CSRC, "synth.c")
    if force or _stale(SYNTH_LIB, [src]):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", SYNTH_LIB, src, "-lm"], check=True)
    return SYNTH_LIB


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_synth(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
