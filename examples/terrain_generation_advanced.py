"""The consumer side of the reference's examples/terrain_generation_advanced.rs through the Python mirror:
noise-driven erodibility, an ocean mask as explicit outlets (:136-210), generate() (:212-216), then the render
loop (:285-315) -- two get_elevation calls per pixel, the second one displaced by the shadow vector -- as two
device rasters instead of 2 x width x height calls.

    python examples/terrain_generation_advanced.py [n_sites] [image_size] [out.npy]

Needs a CUDA device (there is no CPU fallback); writes the brightness-shaded elevation image as float64 .npy
(the PNG colouring of the example is left to the caller).
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastlem_b200 as fl  # noqa: E402
from tools import workloads as W  # noqa: E402


def render(terrain, bound_min, bound_range, img_width, img_height):
    """terrain_generation_advanced.rs:285-315.  Returns (elevation, brightness), NaN where the example skips the pixel."""
    shadow_dist = 0.3
    shadow_angle = 3.14 * 0.25
    shadow_dist_x = shadow_dist * math.cos(shadow_angle)
    shadow_dist_y = shadow_dist * math.sin(shadow_angle)
    shadow_elevation = 50.0
    # x = (bound_range.x - shadow_dist_x) * ((imgx + 0.5) / img_width) + bound_min.x          (:296-299)
    span_x, span_y = bound_range[0] - shadow_dist_x, bound_range[1] - shadow_dist_y
    elevation = terrain.raster(img_width, img_height, bound_min[0], bound_min[1], span_x, span_y, pixel_offset=0.5)
    # site2 = (x + shadow_dist_x, y + shadow_dist_y)                                           (:301-304)
    elevation2 = terrain.raster(img_width, img_height, bound_min[0] + shadow_dist_x, bound_min[1] + shadow_dist_y,
                                span_x, span_y, pixel_offset=0.5)
    # if let (Some(elevation), Some(elevation2)) = ...   brightness = 1 - sin(atan((e - e2) / shadow_elevation))
    both = ~np.isnan(elevation) & ~np.isnan(elevation2)
    brightness = np.full_like(elevation, np.nan)
    brightness[both] = 1.0 - np.sin(np.arctan((elevation[both] - elevation2[both]) / shadow_elevation))
    elevation = np.where(both, elevation, np.nan)
    return elevation, brightness


def main(n_sites=100000, image=512, out=None, lib_path=None):
    bound_min, bound_max = (0.0, 0.0), (100.0, 100.0)
    m = W.delaunay_model(W.random_sites(n_sites, bound_min, bound_max, seed=0), lloyd=1, bound_min=bound_min,
                         bound_max=bound_max)
    p = W.advanced_params(m, seed=0, ocean_level=-0.25)
    params = fl.ParameterArrays(p["base"], p["erodibility"], p["uplift"], p["is_outlet"])
    gen = fl.TerrainGenerator.default().set_model(fl.TerrainModel2D.from_workload(m)).set_parameters(params)
    gen._lib_path = lib_path  # tests run the host emulation build; None = the product library
    terrain = gen.generate()
    elevation, brightness = render(terrain, bound_min, (bound_max[0] - bound_min[0], bound_max[1] - bound_min[1]),
                                   image, image)
    print(f"sites {m['n']} outlets {int(p['is_outlet'].sum())} iterations {gen.last_iterations} "
          f"max elevation {terrain.elevations().max():.3f} pixels inside {np.isfinite(elevation).mean():.3f}")
    if out:
        np.save(out, np.stack([elevation, brightness]))
    return terrain, elevation, brightness


if __name__ == "__main__":
    a = sys.argv[1:]
    main(int(a[0]) if a else 100000, int(a[1]) if len(a) > 1 else 512, a[2] if len(a) > 2 else None)
