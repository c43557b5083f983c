"""Generate tests/golden/*.npz with the independent pure-Python restatement (oracle/pyref.py).

The reference ships no golden vectors for generate() (its tests only write image.png) and cannot be run
here (Rust toolchain absent), so these vectors pin the C++ oracle and the CUDA path to a SECOND restatement
written separately from the first -- not to the crate itself ("parity unpinned").
Re-run:  python tools/make_golden.py     (deterministic; a few seconds)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from tools import workloads as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case(name, m, p, max_iteration=None):
    adj = pyref.adjacency_from_csr(m["row_ptr"], m["col"], m["dist"])
    outlets = [int(o) for o in W.outlets_for(m, p)]
    initial = pyref.initial_elevations([float(b) for b in p["base"]])
    ms = None if p["max_slope"] is None else [float(v) for v in p["max_slope"]]
    k = [float(v) for v in p["erodibility"]]
    u = [float(v) for v in p["uplift"]]
    areas = [float(v) for v in m["areas"]]
    first = pyref.iterate_once(adj, areas, k, u, ms, outlets, initial)
    final, iters = pyref.generate(adj, areas, k, u, ms, outlets, initial, max_iteration)
    none = 0xFFFFFFFF
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        row_ptr=m["row_ptr"].astype(np.uint32), col=m["col"].astype(np.uint32), dist=m["dist"], areas=m["areas"],
        outlets=np.array(outlets, dtype=np.uint32), base=p["base"], erodibility=p["erodibility"], uplift=p["uplift"],
        max_slope=np.array([]) if ms is None else np.array(ms), has_max_slope=np.array(ms is not None),
        max_iteration=np.array(-1 if max_iteration is None else max_iteration),
        initial=np.array(initial), it1_next=np.array(first["next"], dtype=np.uint32),
        it1_next_initial=np.array(first["next_initial"], dtype=np.uint32),
        it1_subroot=np.array(first["subroot"], dtype=np.uint32), it1_has_lake=np.array(first["has_lake"]),
        it1_flood_order=np.array([none if o is None else o for o in first["flood_order"]], dtype=np.uint32),
        it1_drainage=np.array(first["drainage"]), it1_response=np.array(first["response"]),
        it1_elevations=np.array(first["elevations"]), final=np.array(final), iterations=np.array(iters))
    print(f"{name}: n={m['n']} iterations={iters} lakes_in_it1={first['has_lake']}")


def nn_case(name, n_sites, seed, n_queries, bound_max=(100.0, 100.0)):
    """Terrain2D::get_elevation: expected values from the definition-based Sibson restatement (pyref.nn_interpolate)
    at interior queries, plus queries outside the convex hull (expected None = NaN) and on sites (expected = the
    site's value)."""
    m = W.delaunay_model(W.random_sites(n_sites, (0.0, 0.0), bound_max, seed=seed))
    sites, tri, he = W.triangulation_of(m)
    rng = np.random.default_rng(seed + 1)
    values = 50.0 + 40.0 * W.value_noise(sites, 0.05, seed=seed, octaves=3) + rng.random(n_sites)
    bx, by = bound_max
    inner = np.stack([bx * (0.3 + 0.4 * rng.random(n_queries)), by * (0.3 + 0.4 * rng.random(n_queries))], axis=1)
    expected = np.array([pyref.nn_interpolate(sites, values, q)[0] for q in inner])
    outside = np.array([[-1.0, 50.0], [bx + 1.0, 20.0], [30.0, -0.5], [40.0, by + 2.0], [-5.0, -5.0]])
    on_sites = sites[rng.integers(0, n_sites, 5)]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sites=sites, triangles=tri, halfedges=he, values=values,
                        queries=inner, expected=expected, outside=outside, on_sites=on_sites)
    print(f"{name}: n={n_sites} triangles={tri.size // 3} queries={n_queries}")


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--nn-only" not in sys.argv:
        generate_cases()
    nn_case("nn_delaunay150", 150, 201, 40)
    nn_case("nn_delaunay400_wide", 400, 202, 40, bound_max=(200.0, 100.0))


def generate_cases():
    chain = dict(n=4, row_ptr=np.array([0, 1, 3, 5, 6], dtype=np.uint32), col=np.array([1, 0, 2, 1, 3, 2], dtype=np.uint32),
                 dist=np.array([1.0, 1.0, 2.0, 2.0, 0.5, 0.5]), areas=np.array([1.0, 2.0, 3.0, 4.0]),
                 default_outlets=np.array([0], dtype=np.uint32), sites=np.zeros((4, 2)))
    case("chain4", chain, W.uniform_params(4))

    m = W.delaunay_model(W.random_sites(120, seed=101), lloyd=1, bound_min=(0, 0), bound_max=(100, 100))
    case("delaunay120_uniform", m, W.uniform_params(m["n"]))

    m = W.delaunay_model(W.random_sites(400, (0, 0), (200, 100), seed=102), lloyd=1, bound_min=(0, 0), bound_max=(200, 100))
    p = W.uniform_params(m["n"])
    p["max_slope"] = np.full(m["n"], 3.14 * 0.1)
    case("delaunay400_maxslope", m, p)

    m = W.delaunay_model(W.random_sites(300, seed=103))
    p = W.advanced_params(m, seed=4, ocean_level=0.0)
    rng = np.random.default_rng(1)
    ms = 0.1 + rng.random(m["n"]) * 0.5
    ms[rng.random(m["n"]) < 0.5] = np.nan
    p["max_slope"] = ms
    case("delaunay300_advanced_mixed", m, p)

    m = W.delaunay_model(W.random_sites(250, seed=104))
    p = W.uniform_params(m["n"])
    p["uplift"] = 1.1 + 0.9 * W.value_noise(m["sites"], 0.05, seed=3, octaves=2)
    case("delaunay250_uplift", m, p, max_iteration=20)

    m = W.delaunay_model(W.random_sites(150, seed=105))
    p = W.uniform_params(m["n"])
    p["base"] = np.full(m["n"], 2.5)
    case("delaunay150_plateau", m, p)

    m = W.lattice_model(9, 7, jitter=0.0, seed=1)
    case("lattice63_regular", m, W.uniform_params(m["n"]))

    m = W.lattice_model(12, 10, jitter=0.3, seed=2)
    p = W.uniform_params(m["n"])
    p["is_outlet"][[5, 17, 60]] = True
    case("lattice120_interior_outlets", m, p)


if __name__ == "__main__":
    main()
