"""A/B of build-time variants of K1 (k_receivers_mask, fl_paths.cuh: FL_K1_MINBLOCKS, FL_K1_BATCH).
    python tools/ab_k1.py build                 # here: nvcc builds tools/_dbg/libfastlem_b200_k1_<variant>.so
    python tools/ab_k1.py run [sites] [iters]   # on the GPU box: same model, every variant; per-launch time of K1
Results are bit-identical across variants (checked: sha1 of the elevations after `iters` iterations)."""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import build as B  # noqa: E402

DBG = os.path.join(ROOT, "tools", "_dbg")
VARIANTS = {  # the library default is FL_K1_BATCH=3, FL_K1_MINBLOCKS=0
    "default": [],
    "batch8": ["-DFL_K1_BATCH=8"],
    "batch4": ["-DFL_K1_BATCH=4"],
    "batch2": ["-DFL_K1_BATCH=2"],
    "batch3_min6": ["-DFL_K1_MINBLOCKS=6"],
}


def lib_of(name):
    return os.path.join(DBG, f"libfastlem_b200_k1_{name}.so")


def build():
    os.makedirs(DBG, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", lib_of(name)] + B.SOURCES
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out = p.communicate()[0]
        lines = out.splitlines()
        k = next(i for i, ln in enumerate(lines) if "k_receivers_mask" in ln and "Compiling" in ln)
        print(name, "|", lines[k + 2].strip(), "|", lines[k + 3].strip())
        assert p.returncode == 0, out[-2000:]


def run(n, iters):
    from fastlem_b200 import _native
    from tools import workloads as W
    m = W.delaunay_model(W.random_sites(n, seed=1))
    p = W.uniform_params(m["n"])
    outlets = W.outlets_for(m, p)
    initial = _native.host_initial_elevations(p["base"])
    for name in VARIANTS:
        path = lib_of(name)
        if not os.path.exists(path):
            continue
        with _native.Context(0, path) as ctx:
            ctx.set_option("profile", 1)
            ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
            ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, outlets)
            ctx.run(iters)
            it = ctx.run(iters)
            st = ctx.stats()
            e = ctx.download()
        us = 1e3 * st["ms_receivers"] / max(st["n_receivers"], 1)
        gbs = 88.125 * m["n"] / (us * 1e-6) / 1e9
        print(f"{name:14s} sites={m['n']} iterations={it} K1 {us:7.2f} us/launch  {gbs:7.1f} GB/s algorithmic  "
              f"run {st['ms_run']:.1f} ms  sha1 {hashlib.sha1(e.tobytes()).hexdigest()[:12]}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1000000, int(sys.argv[3]) if len(sys.argv) > 3 else 300)
