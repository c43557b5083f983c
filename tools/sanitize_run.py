"""Small end-to-end runs for `compute-sanitizer` (memcheck / initcheck / synccheck / racecheck): three solver scenarios that
together reach every default-path kernel (lake removal in every iteration, the slope clamp, incremental and full K4,
layout rebuilds, the split K5 sweep with a forced run queue) plus one get_elevation raster, each checked against the
oracle.      compute-sanitizer --tool memcheck python tools/sanitize_run.py [sites]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastlem_b200 import _native  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import workloads as W  # noqa: E402
import helpers  # noqa: E402
from scenarios import scenario  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    for name, opts in (("uniform", {}), ("uplift", {}), ("max_slope", {"k5_cut": 1}), ("advanced", {"k5_top_cap": 200})):
        m, p, outlets, initial, max_iteration = scenario(name, n)
        with _native.Context(0) as ctx:
            for k, v in opts.items():
                ctx.set_option(k, v)
            helpers.load_ctx(ctx, m, p, outlets, initial)
            e, it = ctx.generate(max_iteration)
            st = ctx.stats()
        ref, ref_it = O.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, max_iteration)
        assert it == ref_it and np.array_equal(e, ref), name
        print(f"{name}: n={m['n']} iterations={it} lake_iterations={st['lake_iterations']} rebuilds={st['rebuilds']} "
              f"incremental={st['incremental_iterations']} launches={st['kernel_launches']} bit-exact")
    m, p, outlets, initial, _ = scenario("uniform", n)
    sites, tri, he = W.triangulation_of(m)
    with _native.Interpolator(sites, tri, he, device=0) as it:
        it.set_values(e[:sites.shape[0]] if e.size >= sites.shape[0] else np.zeros(sites.shape[0]))
        img = it.raster(it.raster_desc(256, 256, 0.0, 0.0, 100.0, 100.0, 0.5))
    print("raster:", img.shape, float(np.nanmax(img)))


if __name__ == "__main__":
    main()
