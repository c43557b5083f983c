"""A/B: generate() with default options vs the conservative paths (no incremental K4, per-level K5 launches, first
iteration level-synchronous, host flood replay) on a large workload; results must be bit-identical.
   python tools/ab_compare.py SITES [--lattice]"""
import hashlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import _native
from tools import workloads as W
n = int(sys.argv[1]); lattice = "--lattice" in sys.argv
if lattice:
    side = int(round(n ** 0.5)); m = W.lattice_model(side, side, seed=1)
else:
    m = W.delaunay_model(W.random_sites(n, seed=1))
n = m["n"]
p = W.uniform_params(n)
initial = _native.host_initial_elevations(p["base"])
res = []
for opts in ({}, dict(incremental=0, fuse_levels=0, first_flow=0, flood_device=0)):
    with _native.Context(0) as ctx:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, m["default_outlets"])
        t0 = time.perf_counter(); e, it = ctx.generate(); dt = time.perf_counter() - t0
        st = ctx.stats()
        print(f"opts={opts} sites={n} iterations={it} wall={dt:.2f}s device={st['ms_run']/1e3:.2f}s flood_ms={st['ms_flood_rank']:.0f} "
              f"flood_on_device={st['flood_on_device']} incr={st['incremental_iterations']} sha1={hashlib.sha1(e.tobytes()).hexdigest()[:16]}", flush=True)
        res.append((e, it))
print("bit-identical:", res[0][1] == res[1][1] and np.array_equal(res[0][0], res[1][0]))
