"""Run generate() on a synthetic workload and print the library's stats (for ncu / quick experiments).
   python tools/profile_run.py --sites 1000000 --sweep 3 --max-iter 40 [--skip 0] [--lattice]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import _native  # noqa: E402
from tools import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1000000)
    ap.add_argument("--sweep", type=int, default=None)
    ap.add_argument("--rebuild-every", type=int, default=None)
    ap.add_argument("--max-iter", type=int, default=None)
    ap.add_argument("--lattice", action="store_true")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--max-slope", type=float, default=None)
    ap.add_argument("--park-after", type=int, default=None)
    ap.add_argument("--brief", action="store_true")
    ap.add_argument("--lib", default=None, help="path of a variant build of the library (tools/ab_build.py)")
    ap.add_argument("--opt", action="append", default=[], help="solver option name=value (repeatable)")
    args = ap.parse_args()
    cache = f"/tmp/fl_workload_{args.sites}_{int(args.lattice)}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        m = {k: z[k] for k in z.files}
        m["n"] = int(m["row_ptr"].size - 1)
    else:
        if args.lattice:
            side = int(round(args.sites ** 0.5))
            m = W.lattice_model(side, side, seed=1)
        else:
            m = W.delaunay_model(W.random_sites(args.sites, seed=1))
        np.savez(cache, row_ptr=m["row_ptr"], col=m["col"], dist=m["dist"], areas=m["areas"],
                 default_outlets=m["default_outlets"])
    n = m["n"]
    p = W.uniform_params(n)
    initial = _native.host_initial_elevations(p["base"])
    tan = None if args.max_slope is None else np.full(n, np.tan(args.max_slope))
    with _native.Context(0, args.lib) as ctx:
        ctx.set_option("profile", 1)
        if args.sweep is not None:
            ctx.set_option("sweep", args.sweep)
        if args.rebuild_every is not None:
            ctx.set_option("rebuild_every", args.rebuild_every)
        if args.park_after is not None:
            ctx.set_option("park_after", args.park_after)
        for kv in args.opt:
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
        ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctx.set_parameters(initial, p["erodibility"], p["uplift"], tan, m["default_outlets"])
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            it = ctx.run(args.max_iter)
            dt = time.perf_counter() - t0
            st = ctx.stats()
            st["wall_s"] = dt
            st["sites"] = n
            st["msites_per_s_per_iter"] = n * it / dt / 1e6
            st["ms_per_iter"] = 1e3 * dt / max(it, 1)
            if args.brief:
                it_ = max(it, 1)
                print(f"{os.path.basename(args.lib) + ' ' if args.lib else ''}sites={n} opts={args.opt} incr={st['incremental_iterations']} iters={it} ms/iter={st['ms_per_iter']:.3f} K1={st['ms_receivers']/it_:.3f} "
                      f"order={st['ms_order']/it_:.3f} K4={st['ms_area']/it_:.3f} K5={st['ms_elevation']/it_:.3f} rebuilds={st['rebuilds']} "
                      f"levels={st['path_levels']} segs={st['paths']}")
            else:
                print(json.dumps(st))


if __name__ == "__main__":
    main()
