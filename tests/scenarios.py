"""Named test scenarios in the boundary format: each returns
(model, params, outlets, initial_elevation, max_iteration)."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tools import workloads as W  # noqa: E402


def _finish(m, p, max_iteration=None):
    outlets = W.outlets_for(m, p)
    initial = O.initial_elevations(p["base"])
    return m, p, outlets, initial, max_iteration


def _delaunay(n, seed, bound_max=(100.0, 100.0), lloyd=1):
    pts = W.random_sites(n, (0.0, 0.0), bound_max, seed=seed)
    return W.delaunay_model(pts, lloyd=lloyd, bound_min=(0.0, 0.0), bound_max=bound_max)


@functools.lru_cache(maxsize=None)
def scenario(name, n=None):
    if name == "uniform":  # examples/landscape_evolution.rs: random sites, relaxate_sites(1), k = 1, hull outlets
        m = _delaunay(n or 3000, seed=10)
        return _finish(m, W.uniform_params(m["n"]))
    if name == "rust_sites":  # same, but the site stream of builder.rs:38-46 (StdRng::from_seed([0;32]))
        nn = n or 2000
        pts = O.random_sites(nn, (0.0, 0.0), (100.0, 100.0), 0)
        m = W.delaunay_model(pts, lloyd=1, bound_min=(0.0, 0.0), bound_max=(100.0, 100.0))
        return _finish(m, W.uniform_params(nn))
    if name == "max_slope":  # tests/landscape_evolution.rs:19-32: 200x100 box, max_slope = 3.14*0.1 everywhere
        m = _delaunay(n or 2500, seed=11, bound_max=(200.0, 100.0))
        p = W.uniform_params(m["n"])
        p["max_slope"] = np.full(m["n"], 3.14 * 0.1)
        return _finish(m, p)
    if name == "mixed_slope":  # Some(max_slope) on a part of the sites only, varying angle
        m = _delaunay(n or 2000, seed=12)
        p = W.uniform_params(m["n"])
        rng = np.random.default_rng(5)
        ms = 0.05 + rng.random(m["n"]) * 0.6
        ms[rng.random(m["n"]) < 0.5] = np.nan
        p["max_slope"] = ms
        return _finish(m, p)
    if name == "uplift":  # non-uniform uplift: lakes in every iteration, no convergence -> max_iteration
        m = _delaunay(n or 4000, seed=13)
        p = W.uniform_params(m["n"])
        p["uplift"] = 1.1 + 0.9 * W.value_noise(m["sites"], 0.05, seed=3, octaves=2)
        return _finish(m, p, max_iteration=25)
    if name == "advanced":  # terrain_generation_advanced.rs style: noise erodibility, ocean-mask outlets
        m = _delaunay(n or 5000, seed=14)
        p = W.advanced_params(m, seed=2, ocean_level=-0.1)
        return _finish(m, p)
    if name == "plateau":  # base_elevation != 0 absorbs the noise: everything is a plateau of lakes
        m = _delaunay(n or 1500, seed=15)
        p = W.uniform_params(m["n"])
        p["base"] = np.full(m["n"], 3.0)
        return _finish(m, p)
    if name == "base_field":  # smooth non-zero base elevation field
        m = _delaunay(n or 2000, seed=16)
        p = W.uniform_params(m["n"])
        p["base"] = 2.0 + W.value_noise(m["sites"], 0.08, seed=9, octaves=3)
        return _finish(m, p)
    if name in ("edge_sites_ocean", "edge_sites_partial"):
        # terrain_generation_advanced.rs:36-42,178-182: relaxed random sites + add_edge_sites(None, None) (equally spaced
        # rim sites), outlets = ocean flood-filled from the rim.
        # _ocean: the whole rim is ocean (every tied edge joins two outlets); _partial: parts of the rim are land
        nn = n or 3000
        pts = W.random_sites(nn, seed=23)
        from scipy.spatial import Delaunay
        pts = W._lloyd_step(pts, Delaunay(pts), (0.0, 0.0), (100.0, 100.0))
        # 64 sites per side: spacing 100/64 is exact in binary, so ALL rim edges have the same length bit for bit (with the
        # default count the lerp's rounding makes most of them differ in the last place)
        m = W.delaunay_model_with_rim(W.add_edge_sites(pts, edge_num_x=64, edge_num_y=64), nn)
        p = W.uniform_params(m["n"])
        p["erodibility"] = np.abs(W.value_noise(m["sites"], 8.0 / 75.0, seed=4, octaves=3)) * 4.0 + 0.1
        p["is_outlet"] = W.ocean_rim_outlets(m, nn, band=3.0 if name == "edge_sites_ocean" else None, seed=5)
        assert p["is_outlet"].any()
        return _finish(m, p)
    if name == "lattice":  # equal rim edge lengths: exact key ties in the flood heap
        m = W.lattice_model(40, 30, jitter=0.3, seed=4)
        return _finish(m, W.uniform_params(m["n"]))
    if name == "lattice_regular":  # no jitter at all: ties everywhere (slopes, edge lengths)
        m = W.lattice_model(24, 20, jitter=0.0, seed=5)
        return _finish(m, W.uniform_params(m["n"]))
    if name == "single_outlet":
        m = _delaunay(n or 1200, seed=17)
        p = W.uniform_params(m["n"])
        p["is_outlet"][m["default_outlets"][0]] = True
        return _finish(m, p)
    if name == "interior_outlets":  # outlets in the interior only (hull sites become ordinary nodes)
        m = _delaunay(n or 1500, seed=18)
        p = W.uniform_params(m["n"])
        rng = np.random.default_rng(7)
        p["is_outlet"][rng.choice(m["n"], 12, replace=False)] = True
        return _finish(m, p)
    if name == "disconnected":  # second component without any outlet: never visited (generator.rs:149)
        a = _delaunay(700, seed=19)
        b = _delaunay(300, seed=20)
        na = a["n"]
        m = dict(n=na + b["n"],
                 row_ptr=np.concatenate([a["row_ptr"], b["row_ptr"][1:] + a["row_ptr"][-1]]).astype(np.uint32),
                 col=np.concatenate([a["col"], b["col"] + na]).astype(np.uint32),
                 dist=np.concatenate([a["dist"], b["dist"]]),
                 areas=np.concatenate([a["areas"], b["areas"]]),
                 default_outlets=a["default_outlets"],
                 sites=np.concatenate([a["sites"], b["sites"] + 200.0]))
        return _finish(m, W.uniform_params(m["n"]))
    if name == "hub":  # one site adjacent to 60 others: rows longer than the 32-bit child mask
        base = _delaunay(n or 900, seed=21)
        nn = base["n"]
        rp = base["row_ptr"].astype(np.int64)
        rows = [list(zip(base["col"][rp[i]:rp[i + 1]].tolist(), base["dist"][rp[i]:rp[i + 1]].tolist()))
                for i in range(nn)]
        hub = nn // 2
        rng = np.random.default_rng(3)
        have = {j for j, _ in rows[hub]} | {hub}
        extra = [int(j) for j in rng.permutation(nn) if int(j) not in have][:60]
        for j in extra:  # add_edge(hub, j, w): append to both lists
            w = float(np.hypot(*(base["sites"][hub] - base["sites"][j])))
            rows[hub].append((j, w))
            rows[j].append((hub, w))
        m = dict(base)
        m["row_ptr"] = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint32)
        m["col"] = np.array([j for r in rows for j, _ in r], dtype=np.uint32)
        m["dist"] = np.array([w for r in rows for _, w in r], dtype=np.float64)
        return _finish(m, W.uniform_params(nn))
    if name == "tiny_chain":  # 0 - 1 - 2 - 3, outlet 0
        m = dict(n=4, row_ptr=np.array([0, 1, 3, 5, 6], dtype=np.uint32),
                 col=np.array([1, 0, 2, 1, 3, 2], dtype=np.uint32),
                 dist=np.array([1.0, 1.0, 2.0, 2.0, 0.5, 0.5]), areas=np.array([1.0, 2.0, 3.0, 4.0]),
                 default_outlets=np.array([0], dtype=np.uint32), sites=np.zeros((4, 2)))
        return _finish(m, W.uniform_params(4))
    if name == "isolated_nodes":  # sites without any edge (builder.rs:254-266 can leave a hull site edgeless)
        m = dict(n=3, row_ptr=np.array([0, 1, 2, 2], dtype=np.uint32), col=np.array([1, 0], dtype=np.uint32),
                 dist=np.array([1.5, 1.5]), areas=np.array([1.0, 1.0, 1.0]),
                 default_outlets=np.array([0], dtype=np.uint32), sites=np.zeros((3, 2)))
        return _finish(m, W.uniform_params(3))
    raise KeyError(name)


SMALL = ["uniform", "rust_sites", "max_slope", "mixed_slope", "uplift", "advanced", "plateau", "base_field",
         "lattice", "lattice_regular", "single_outlet", "interior_outlets", "disconnected", "hub", "tiny_chain",
         "isolated_nodes"]
