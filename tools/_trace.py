import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastlem_b200 import _native
from tools import workloads as W
n = int(sys.argv[1])
cache = f"/tmp/fl_workload_{n}_0.npz"
z = np.load(cache); m = {k: z[k] for k in z.files}; m["n"] = n
p = W.uniform_params(n)
initial = _native.host_initial_elevations(p["base"])
for rep in range(2):
    t0 = time.perf_counter()
    ctx = _native.Context(0)
    ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
    ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, m["default_outlets"])
    t1 = time.perf_counter()
    e, it = ctx.generate(3)
    t2 = time.perf_counter()
    st = ctx.stats()
    ctx.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: setup {t1-t0:.3f} generate(3) {t2-t1:.3f} close {t3-t2:.3f} flood_ms {st['ms_flood_rank']:.1f} on_device {st['flood_on_device']}", flush=True)
