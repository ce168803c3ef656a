"""In-tree build of libeals_b200.so (nvcc, sm_100a only).  Cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libeals_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "eals_b200.h"))
    return deps


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/eals_b200.cu -> lib/libeals_b200.so if any source is newer than the library."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB):
        t = os.path.getmtime(LIB)
        if all(os.path.getmtime(s) <= t for s in _sources()):
            return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "eals_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB


def build_host_example(force: bool = False) -> str:
    """Compile the C++ drop-in driver (host/eals_main.cpp over include/MF_fastALS.h) against the
    library — the 'same three call sites' check of SURVEY.md §8b."""
    src = os.path.join(HERE, "host", "eals_main.cpp")
    out = os.path.join(LIB_DIR, "eals_main")
    if not os.path.exists(src):
        return ""
    deps = [src, os.path.join(ROOT, "include", "MF_fastALS.h"), os.path.join(ROOT, "include", "eals_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                    "-L", LIB_DIR, "-leals_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_host_example(force=True))
