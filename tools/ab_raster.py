"""A/B of build-time variants of the query kernel (fl_interp.cuh: FLI_MINBLOCKS, FLI_RECIP, FLI_UNIFIED, FLI_WARP_8X4).
    python tools/ab_raster.py build            # here: nvcc builds tools/_dbg/libfastlem_b200_<variant>.so
    python tools/ab_raster.py check            # here: the same variants as host emulation builds against the oracle
    python tools/ab_raster.py run [n] [size]   # on the GPU box: one model, every variant, kernel times by CUDA events
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import build as B  # noqa: E402

DBG = os.path.join(ROOT, "tools", "_dbg")
_OFF = ["-DFLI_MINBLOCKS=0", "-DFLI_RECIP=0", "-DFLI_UNIFIED=0", "-DFLI_WARP_8X4=0", "-DFLI_DENORM=0"]
VARIANTS = {  # the library's defaults are all on (MINBLOCKS=4); each variant overrides some of them
    "default": [],
    "all_off": _OFF,
    "no_denorm": ["-DFLI_DENORM=0"],
    "min3": ["-DFLI_MINBLOCKS=3"],
    "min5": ["-DFLI_MINBLOCKS=5"],
    "no_recip": ["-DFLI_RECIP=0"],
    "no_unified": ["-DFLI_UNIFIED=0"],
    "w16x2": ["-DFLI_WARP_8X4=0"],
}


def lib_of(name, emu=False):
    return os.path.join(DBG, f"libfastlem_{'emu' if emu else 'b200'}_{name}.so")


def build():
    os.makedirs(DBG, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", lib_of(name)] + B.SOURCES
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out = p.communicate()[0]
        lines = out.splitlines()
        k = next(i for i, ln in enumerate(lines) if "k_nn_raster" in ln and "Compiling" in ln)
        print(name, "|", lines[k + 2].strip(), "|", lines[k + 3].strip())
        assert p.returncode == 0, out[-2000:]


def check():
    from fastlem_b200 import _native
    from oracle import oracle as O
    from tools import workloads as W
    os.makedirs(DBG, exist_ok=True)
    m = W.delaunay_model(W.random_sites(3000, seed=5))
    sites, tri, he = W.triangulation_of(m)
    values = 20.0 + 10.0 * W.value_noise(sites, 0.07, seed=2, octaves=3)
    cols, rows = np.meshgrid(np.arange(96.0), np.arange(96.0))
    q = np.stack([100.0 * ((cols.reshape(-1) + 0.5) / 96), 100.0 * ((rows.reshape(-1) + 0.5) / 96)], axis=1)
    ref = O.nn_interpolate(sites, tri, values, q).reshape(96, 96)
    for name, flags in VARIANTS.items():
        cmd = ["g++", "-O2", "-std=c++17", "-DFL_EMU", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-x",
               "c++"] + flags + B.SOURCES + ["-o", lib_of(name, True)]
        subprocess.check_call(cmd)
        with _native.Interpolator(sites, tri, he, lib_path=lib_of(name, True)) as it:
            it.set_values(values)
            img = it.raster(it.raster_desc(96, 96, 0.0, 0.0, 100.0, 100.0, 0.5))
        assert np.array_equal(np.isnan(img), np.isnan(ref))
        ok = ~np.isnan(ref)
        print(name, "max rel err vs oracle", float((np.abs(img[ok] - ref[ok]) / np.maximum(1, np.abs(ref[ok]))).max()))


def run(n, size):
    from fastlem_b200 import _native
    from tools import workloads as W
    m = W.delaunay_model(W.random_sites(n, seed=1))
    sites, tri, he = W.triangulation_of(m)
    values = 0.5 * sites[:, 0] + 0.25 * sites[:, 1] - 3.0
    cols, rows = np.meshgrid(np.arange(size, dtype=np.float64), np.arange(size, dtype=np.float64))
    plane = 0.5 * (100.0 * ((cols + 0.5) / size)) + 0.25 * (100.0 * ((rows + 0.5) / size)) - 3.0
    for name in VARIANTS:
        path = lib_of(name)
        if not os.path.exists(path):
            continue
        with _native.Interpolator(sites, tri, he, lib_path=path) as it:
            it.set_values(values)
            desc = it.raster_desc(size, size, 0.0, 0.0, 100.0, 100.0, 0.5)
            ms = []
            for _ in range(6):
                img = it.raster(desc)
                ms.append(it.stats()["ms_query_kernel"])
        ok = ~np.isnan(img)
        print(f"{name:20s} sites={n} raster={size}^2 kernel_ms min={min(ms[1:]):.3f} median={np.median(ms[1:]):.3f} "
              f"Gpixel/s={size * size / min(ms[1:]) / 1e6:.2f} plane_err={np.abs(img[ok] - plane[ok]).max():.2e} "
              f"inside={ok.mean():.4f}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "run"
    if what == "build":
        build()
    elif what == "check":
        check()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 250000, int(sys.argv[3]) if len(sys.argv) > 3 else 2048)
