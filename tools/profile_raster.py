"""One get_elevation raster on cuda:0 for ncu captures of k_nn_raster (profiles/): 
    ncu --set full --clock-control none --import-source on -k k_nn_raster -c 1 -o gpurun_out/raster python tools/profile_raster.py
Prints the CUDA-event time of the kernel when run without a profiler."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import _native  # noqa: E402
from tools import workloads as W  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    m = W.delaunay_model(W.random_sites(n, seed=1))
    sites, tri, he = W.triangulation_of(m)
    values = 10.0 * W.value_noise(sites, 0.05, seed=1, octaves=4)
    with _native.Interpolator(sites, tri, he) as it:
        it.set_values(values)
        desc = it.raster_desc(size, size, 0.0, 0.0, 100.0, 100.0, 0.5)
        for _ in range(reps):
            img = it.raster(desc)
            st = it.stats()
            print(f"sites={n} raster={size}x{size} kernel_ms={st['ms_query_kernel']:.3f} "
                  f"Mpixel/s={size * size / st['ms_query_kernel'] / 1e3:.1f} inside={np.isfinite(img).mean():.3f} "
                  f"setup_ms={st['ms_setup']:.1f} grid={st['grid_x']}x{st['grid_y']} passes={st['grid_passes']}")


if __name__ == "__main__":
    main()
