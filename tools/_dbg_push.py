import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastlem_b200 import _native
from tools import workloads as W
for n in [int(a) for a in sys.argv[1:]] or (500, 3000, 30000):
    m = W.delaunay_model(W.random_sites(n, seed=1))
    p = W.uniform_params(m["n"])
    outlets = W.outlets_for(m, p)
    initial = _native.host_initial_elevations(p["base"])
    for mi in (1, 3, None):
        try:
            with _native.Context(0) as ctx:
                ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
                ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, outlets)
                t=time.time(); it = ctx.run(mi); print(n, mi, "ok", it, time.time()-t, flush=True)
        except Exception as ex:
            print(n, mi, "FAIL", ex, flush=True)
