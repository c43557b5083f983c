"""Event counters of the dataflow K4 (needs tools/_dbg/libfastlem_stats.so built with -DFL_FLOW_STATS)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import _native
from tools import workloads as W
LIB = os.path.join(ROOT, "tools", "_dbg", "libfastlem_stats.so")
NAMES = ["T_CLIMBS", "T_BATCH", "T_SITES", "T_HEADS", "T_LAST", "T_SEGSTART", "T_PARKED", "T_NOTREADY",
         "W_FLOWS", "W_WINDOWS", "W_SITES", "W_HEADS", "W_LAST", "W_SEGSTART", "W_REDO", "W_NOTREADY",
         "W_CYC_FLOW", "W_CYC_REPORT", "W_CYC_FIRSTWIN", "W_CYC_WIN"]

def fetch(ctx):
    out = np.zeros(32, dtype=np.uint64)
    ctx._ck(ctx._lib.fastlem_debug_fetch(ctx._h, 9, out.ctypes.data_as(ctypes.c_void_p), 256))
    return out

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cache = f"/tmp/fl_workload_{n}_0.npz"
if os.path.exists(cache):
    z = np.load(cache); m = {k: z[k] for k in z.files}; m["n"] = n
else:
    m = W.delaunay_model(W.random_sites(n, seed=1))
p = W.uniform_params(n)
initial = _native.host_initial_elevations(p["base"], LIB)
with _native.Context(0, LIB) as ctx:
    for k, v in [a.split("=") for a in sys.argv[2:]]:
        ctx.set_option(k, int(v))
    ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
    ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, m["default_outlets"])
    ctx.run(100); a = fetch(ctx)
    ctx.run(300); b = fetch(ctx)
    d = (b.astype(np.int64) - a.astype(np.int64)) / 200.0
    print("per iteration, iterations 101-300:")
    for k, nm in enumerate(NAMES):
        print(f"  {nm:10s} {d[k]:12.1f}")
