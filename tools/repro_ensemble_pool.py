"""4M-site Delaunay ensemble: the pool of bench.py's multi-GPU arm on ONE GPU (4 contexts, context reuse, to convergence)."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fastlem_b200 import _native, ensemble
t0 = time.time()
m, p, outlets, _ = bench.build_workload("delaunay", int(sys.argv[1]) if len(sys.argv) > 1 else 4000000, 1)
n = m["n"]
basis = bench.erodibility_basis(m["sites"])
initial = _native.host_initial_elevations(p["base"])
print("built", n, outlets.size, time.time() - t0, flush=True)
def run(lib, n_ctx, members):
    ctxs = []
    for k in range(n_ctx):
        c = _native.Context(0, lib)
        c.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctxs.append(c)
    pool = ensemble.MemberPool(members)
    lock = threading.Lock()
    log = []
    def on_result(t, it, ctx):
        st = ctx.stats()
        with lock:
            log.append((t, it, round(st["ms_run"])))
    def make_params(t):
        return dict(initial=initial, erodibility=bench.member_erodibility(basis, t), uplift=p["uplift"], outlets=outlets)
    t1 = time.time()
    try:
        ensemble.run_pool_concurrent(ctxs, pool, make_params, on_result)
        msg = "ok"
    except Exception as ex:
        msg = f"FAILED {ex}"
    print(os.path.basename(lib or "default"), n_ctx, "contexts", members, "members", round(time.time() - t1, 2), "s", msg, sorted(log), flush=True)
    for c in ctxs:
        try: c.close()
        except Exception: pass
D = os.path.join(ROOT, "tools", "_dbg")
run(None, 4, 16)
run(os.path.join(D, "libfastlem_b200_spin30.so"), 4, 16)
