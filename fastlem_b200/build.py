"""Build the native library in-tree.

  python -m fastlem_b200.build          -> fastlem_b200/_lib/libfastlem_b200.so   (nvcc, sm_100a; the product)
  python -m fastlem_b200.build --emu    -> tests/_emu/libfastlem_emu.so           (g++ -DFL_EMU; CPU test tier only)

nvcc cross-compiles without a GPU.  -fmad=false keeps every double-precision expression at the
reference's rounding points (Rust never contracts a*b+c into an FMA).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = [os.path.join(CSRC, f) for f in ("fl_solver.cu", "fl_interp.cu", "fl_flood.cpp", "fl_host.cpp")]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))) + \
          [os.path.join(ROOT, "include", "fastlem_b200.h")]
LIB = os.path.join(HERE, "_lib", "libfastlem_b200.so")
EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libfastlem_emu.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not _stale(LIB):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd)
    return LIB


def build_emu(force=False):
    if not force and not _stale(EMU_LIB):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-DFL_EMU", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-x", "c++"] + SOURCES + ["-o", EMU_LIB]
    subprocess.check_call(cmd)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force=True))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
