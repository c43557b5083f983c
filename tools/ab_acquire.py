"""A/B of the consumer-side acquire fences of the dataflow hand-offs (fl_flow.cuh, FL_RELAXED_READERS).
    python tools/ab_acquire.py build                    # here: tools/_dbg/libfastlem_b200_acq_<variant>.so
    python tools/ab_acquire.py run [sites] [repeats]    # on the GPU box: same model, generate() to convergence
`acquire` is the shipped build (fence.acq_rel.gpu / ld.acquire.gpu on the consumer side of every hand-off);
`relaxed` is the round-1 code (address dependency on the atomic's result only) kept for timing comparison."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import build as B  # noqa: E402

DBG = os.path.join(ROOT, "tools", "_dbg")
VARIANTS = {"acquire": [], "relaxed": ["-DFL_RELAXED_READERS=1"]}


def lib_of(name):
    return os.path.join(DBG, f"libfastlem_b200_acq_{name}.so")


def build():
    os.makedirs(DBG, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + flags + ["-o", lib_of(name)] + B.SOURCES
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out = p.communicate()[0]
        assert p.returncode == 0, out[-2000:]
        print("built", lib_of(name))


def run(n, repeats):
    from fastlem_b200 import _native
    from tools import workloads as W
    m = W.delaunay_model(W.random_sites(n, seed=1))
    p = W.uniform_params(m["n"])
    outlets = W.outlets_for(m, p)
    initial = _native.host_initial_elevations(p["base"])
    for rep in range(repeats):
        for name in VARIANTS:
            path = lib_of(name)
            if not os.path.exists(path):
                continue
            with _native.Context(0, path) as ctx:
                ctx.set_option("profile", 1)
                ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
                ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, outlets)
                ctx.run(None)
                it = ctx.run(None)
                st = ctx.stats()
                e = ctx.download()
            print(f"{name:8s} sites={m['n']} iterations={it} run {st['ms_run']:.1f} ms  per iteration: K1 "
                  f"{st['ms_receivers'] / it:.4f} order {st['ms_order'] / it:.4f} K4 {st['ms_area'] / it:.4f} "
                  f"K5 {st['ms_elevation'] / it:.4f} ms  sha1 {hashlib.sha1(e.tobytes()).hexdigest()[:12]}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1000000, int(sys.argv[3]) if len(sys.argv) > 3 else 2)
