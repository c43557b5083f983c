"""Terrain2D::get_elevation (natural-neighbour interpolation; reference src/models/surface/terrain.rs:36-38,
src/models/surface/interpolator.rs:6-28; the arithmetic lives in the un-vendored crate naturalneighbor 1.2.2, so
parity is tolerance-based and UNPINNED -- see oracle/nn_oracle.cpp).

CPU tier: the C++ oracle against the golden vectors of the definition-based Python restatement and against the
interpolant's defining properties; the device kernels in their host emulation build against the oracle.
GPU tier (-m gpu): the same comparisons through the product library on cuda:0.
"""
import numpy as np
import pytest

import helpers
from tools import workloads as W

# floating point: relative to max(1, |z|) of values of order 1e2.  The three formulations (device: per-edge shoelace
# terms about (p+v)/2 in double; oracle: ordered polygons in long double; pyref: half-plane clipping) agree to ~1e-13.
NN_TOL = 1e-9


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def make_case(n, seed, bound_max=(100.0, 100.0), lloyd=0):
    m = W.delaunay_model(W.random_sites(n, (0.0, 0.0), bound_max, seed=seed), lloyd=lloyd, bound_min=(0.0, 0.0),
                         bound_max=bound_max)
    sites, tri, he = W.triangulation_of(m)
    rng = np.random.default_rng(seed + 7)
    values = 30.0 + 25.0 * W.value_noise(sites, 0.08, seed=seed, octaves=3) + rng.random(sites.shape[0])
    return m, sites, tri, he, values


def queries_for(bound_max, nq, seed, margin=0.03):
    rng = np.random.default_rng(seed)
    bx, by = bound_max
    return np.stack([bx * (-margin + (1 + 2 * margin) * rng.random(nq)), by * (-margin + (1 + 2 * margin) * rng.random(nq))], 1)


# ------------------------------------------------------------------------------------------------
# oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", helpers.nn_golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_golden(oracle, path):
    g = np.load(path)
    out = oracle.nn_interpolate(g["sites"], g["triangles"], g["values"], g["queries"])
    assert rel_err(out, g["expected"]).max() <= NN_TOL
    assert np.isnan(oracle.nn_interpolate(g["sites"], g["triangles"], g["values"], g["outside"])).all()
    on = oracle.nn_interpolate(g["sites"], g["triangles"], g["values"], g["on_sites"])
    idx = [int(np.nonzero((g["sites"] == q).all(axis=1))[0][0]) for q in g["on_sites"]]
    assert np.array_equal(on, g["values"][idx])


def test_oracle_weights_are_sibson_coordinates(oracle):
    """Partition of unity, positivity, and the local-coordinates property sum_i w_i s_i = p."""
    _, sites, tri, _, _ = make_case(500, 11)
    for q in queries_for((100.0, 100.0), 60, 5, margin=-0.2):
        ids, w = oracle.nn_weights(sites, tri, q[0], q[1])
        assert ids.size >= 3
        assert (w > 0).all() and abs(w.sum() - 1.0) < 1e-12
        assert np.abs((w[:, None] * sites[ids]).sum(axis=0) - q).max() < 1e-10


def test_oracle_walk_location_gives_the_same_values(oracle):
    """The walk-located variant timed by bench.py's CPU baseline is the same interpolation."""
    _, sites, tri, _, values = make_case(2000, 13)
    cols, rows = np.meshgrid(np.arange(64.0), np.arange(8.0))
    q = np.stack([100.0 * cols.reshape(-1) / 64.0, 40.0 + rows.reshape(-1)], axis=1)
    a = oracle.nn_interpolate(sites, tri, values, q)
    b = oracle.nn_interpolate(sites, tri, values, q, walk=True)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and (~np.isnan(a)).sum() > 400
    ok = ~np.isnan(a)
    assert rel_err(b[ok], a[ok]).max() <= 1e-12


def test_pyref_definition_matches_oracle(oracle):
    from oracle import pyref
    _, sites, tri, _, values = make_case(120, 12)
    for q in queries_for((100.0, 100.0), 10, 6, margin=-0.3):
        z, w = pyref.nn_interpolate(sites, values, q)
        ids, ww = oracle.nn_weights(sites, tri, q[0], q[1])
        assert sorted(w) == sorted(int(i) for i in ids)
        assert max(abs(w[int(i)] - x) for i, x in zip(ids, ww)) < 1e-11
        assert abs(z - oracle.nn_interpolate(sites, tri, values, q[None])[0]) <= NN_TOL * max(1.0, abs(z))


# ------------------------------------------------------------------------------------------------
# the device path: shared checks, run on the emulation build (CPU tier) and on the product (GPU tier)
# ------------------------------------------------------------------------------------------------
def check_golden(lib, path):
    from fastlem_b200 import _native
    g = np.load(path)
    with _native.Interpolator(g["sites"], g["triangles"], g["halfedges"], lib_path=lib) as it:
        it.set_values(g["values"])
        assert rel_err(it.points(g["queries"]), g["expected"]).max() <= NN_TOL
        assert np.isnan(it.points(g["outside"])).all()
        idx = [int(np.nonzero((g["sites"] == q).all(axis=1))[0][0]) for q in g["on_sites"]]
        assert np.array_equal(it.points(g["on_sites"]), g["values"][idx])


def check_against_oracle(lib, O, n, seed, nq, bound_max=(100.0, 100.0), lloyd=0):
    from fastlem_b200 import _native
    _, sites, tri, he, values = make_case(n, seed, bound_max, lloyd)
    q = queries_for(bound_max, nq, seed + 1)
    ref = O.nn_interpolate(sites, tri, values, q)
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        it.set_values(values)
        out = it.points(q)
        st = it.stats()
    assert np.array_equal(np.isnan(out), np.isnan(ref)), "None pattern (outside the convex hull)"
    assert np.isnan(ref).any() and (~np.isnan(ref)).sum() > (nq // 2 if n >= 50 else 50)
    ok = ~np.isnan(ref)
    assert rel_err(out[ok], ref[ok]).max() <= NN_TOL
    assert st["queries"] == nq and st["kernel_launches"] >= 4
    return out


def check_properties(lib):
    """Linear precision (a plane is reproduced), bounds (convex combination), values at the sites, hull edges."""
    from fastlem_b200 import _native
    m, sites, tri, he, values = make_case(3000, 21)
    q = queries_for((100.0, 100.0), 5000, 22, margin=-0.05)
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        it.set_values(2.0 * sites[:, 0] - 3.0 * sites[:, 1] + 1.0)
        out = it.points(q)
        assert not np.isnan(out).any()
        assert np.abs(out - (2.0 * q[:, 0] - 3.0 * q[:, 1] + 1.0)).max() < 1e-9
        it.set_values(values)
        out = it.points(q)
        assert (out >= values.min() - 1e-9).all() and (out <= values.max() + 1e-9).all()
        assert np.array_equal(it.points(sites), values)  # p on a site: that site's value
        # a point on a hull edge: linear along the edge
        hull = m["default_outlets"]
        rp, col = m["row_ptr"], m["col"]
        a = int(hull[0])
        nb = col[rp[a]:rp[a + 1]]
        b = int([v for v in nb if v in set(hull.tolist())][0])
        mid = 0.5 * (sites[a] + sites[b])
        z = it.points(mid[None])[0]
        if not np.isnan(z):  # the midpoint may round to just outside
            assert abs(z - 0.5 * (values[a] + values[b])) < 1e-6 * max(1.0, abs(z))
        # NaN query coordinates -> None
        assert np.isnan(it.points(np.array([[np.nan, 1.0], [1.0, np.nan]]))).all()


def check_queries_on_shared_edges(lib, O):
    """Queries on an interior triangle edge within rounding (midpoints and other points of every edge): the two
    triangles evaluate the shared edge from different base vertices, so both can see a tiny negative orientation --
    the visibility walk must not bounce between them (round-1 advisor finding: one such pixel failed the whole call
    with E_INVALID).  The reference answers these points like any other."""
    from fastlem_b200 import _native
    for n, seed in ((300, 0), (300, 1), (2000, 2)):
        rng = np.random.default_rng(seed)
        sites = rng.random((n, 2)) * 100.0
        m = W.delaunay_model(sites)
        sites, tri, he = W.triangulation_of(m)
        values = 30.0 + 25.0 * W.value_noise(sites, 0.08, seed=seed, octaves=3) + rng.random(n)
        t3 = np.asarray(tri).reshape(-1, 3)
        interior = np.asarray(he).reshape(-1, 3) != 0xFFFFFFFF  # delaunator: no opposite half-edge on the hull
        qs = []
        for k in range(3):
            a, b = sites[t3[:, k]][interior[:, k]], sites[t3[:, (k + 1) % 3]][interior[:, k]]
            for s in (0.5, 0.25, 1.0 / 3.0, 0.9):
                qs.append(a + s * (b - a))
        q = np.concatenate(qs)
        with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
            it.set_values(values)
            out = it.points(q)  # raises on a walk overflow
        assert not np.isnan(out).any()
        sub = rng.choice(q.shape[0], 400, replace=False)
        ref = O.nn_interpolate(sites, tri, values, q[sub])
        assert not np.isnan(ref).any()
        assert rel_err(out[sub], ref).max() <= NN_TOL


def check_raster(lib, O):
    """The raster entry point equals per-pixel point queries with the examples' coordinate formula, for both pixel
    offsets, any row block, and is independent of how the rows are partitioned."""
    from fastlem_b200 import _native
    _, sites, tri, he, values = make_case(800, 31, bound_max=(200.0, 100.0))
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        it.set_values(values)
        for width, height, off, x0, y0, sx, sy in ((37, 23, 0.0, 0.0, 0.0, 200.0, 100.0),
                                                    (50, 41, 0.5, 3.0, -2.0, 190.0, 104.0)):
            full = it.raster(it.raster_desc(width, height, x0, y0, sx, sy, off))
            assert full.shape == (height, width)
            cols, rows = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
            px = sx * ((cols + off) / width) + x0   # examples/landscape_evolution.rs:49-50 / advanced.rs:296-299
            py = sy * ((rows + off) / height) + y0
            q = np.stack([px.reshape(-1), py.reshape(-1)], axis=1)
            pts = it.points(q).reshape(height, width)
            assert np.array_equal(full, pts, equal_nan=True)
            ref = O.nn_interpolate(sites, tri, values, q).reshape(height, width)
            assert np.array_equal(np.isnan(full), np.isnan(ref))
            ok = ~np.isnan(ref)
            assert ok.sum() > ok.size // 2
            assert rel_err(full[ok], ref[ok]).max() <= NN_TOL
            parts = [it.raster(it.raster_desc(width, height, x0, y0, sx, sy, off, r0, r1))
                     for r0, r1 in ((0, 7), (7, 8), (8, 8), (8, height))]
            assert parts[2].shape == (0, width)
            assert np.array_equal(np.concatenate(parts, axis=0), full, equal_nan=True)


def check_orientation_and_validation(lib):
    from fastlem_b200 import _native
    _, sites, tri, he, values = make_case(300, 41)
    q = queries_for((100.0, 100.0), 300, 42)
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        it.set_values(values)
        ccw = it.points(q)
        assert it.stats()["clockwise"] == 0
    # the same triangulation clockwise (vertices 1 and 2 swapped; half-edge k of the flipped triangle is old 2-k)
    T = tri.reshape(-1, 3)
    tri_cw = T[:, [0, 2, 1]].reshape(-1)
    he_old = he.reshape(-1, 3)[:, [2, 1, 0]].reshape(-1)
    valid = he_old != 0xFFFFFFFF
    he_cw = he_old.copy()
    he_cw[valid] = (he_old[valid] // 3) * 3 + (2 - he_old[valid] % 3)
    with _native.Interpolator(sites, tri_cw, he_cw, lib_path=lib) as it:
        it.set_values(values)
        cw = it.points(q)
        assert it.stats()["clockwise"] == 1
    assert np.array_equal(np.isnan(cw), np.isnan(ccw))
    ok = ~np.isnan(ccw)
    assert rel_err(cw[ok], ccw[ok]).max() <= NN_TOL
    # rejected inputs
    bad = tri.copy()
    bad[5] = sites.shape[0] + 3
    with pytest.raises(_native.FastlemError):
        _native.Interpolator(sites, bad, he, lib_path=lib)
    bad_he = he.copy()
    i = int(np.nonzero(he != 0xFFFFFFFF)[0][0])
    bad_he[i] = (bad_he[i] + 1) % he.size
    with pytest.raises(_native.FastlemError):
        _native.Interpolator(sites, tri, bad_he, lib_path=lib)
    mixed = T.copy()
    mixed[0] = mixed[0, [0, 2, 1]]
    with pytest.raises(_native.FastlemError):
        _native.Interpolator(sites, mixed.reshape(-1), he, lib_path=lib)
    # a valid but non-Delaunay triangulation (jittered lattice with random diagonals)
    lm = W.lattice_model(12, 12, jitter=0.3, seed=3)
    ls, lt, lh = W.triangulation_of(lm)
    with pytest.raises(_native.FastlemError):
        _native.Interpolator(ls, lt, lh, lib_path=lib)
    # values must be set before a query
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        with pytest.raises(_native.FastlemError) as e:
            it.points(q)
        assert e.value.code == _native.E_STATE
        with pytest.raises(_native.FastlemError):
            it.raster(it.raster_desc(4, 4, 0, 0, 1, 1, 0.0, 3, 9))


def check_values_from_solver(lib, O):
    """Terrain2D built straight from a generate() result: the solver's elevations reach the interpolator on the
    device (fastlem_interp_set_values_from) and give the same raster as a host round trip."""
    from fastlem_b200 import _native
    m = W.delaunay_model(W.random_sites(1500, seed=51), lloyd=1, bound_min=(0, 0), bound_max=(100, 100))
    p = W.uniform_params(m["n"])
    outlets = W.outlets_for(m, p)
    initial = O.initial_elevations(p["base"])
    sites, tri, he = W.triangulation_of(m)
    with _native.Context(0, lib) as ctx, _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        elev, _ = ctx.generate()
        it.set_values_from(ctx)
        desc = it.raster_desc(64, 64, 0.0, 0.0, 100.0, 100.0)
        a = it.raster(desc)
        it.set_values(elev)
        b = it.raster(desc)
    assert np.array_equal(a, b, equal_nan=True)
    assert (~np.isnan(a)).sum() > 3000
    ref = O.nn_interpolate(sites, tri, elev, np.array([[50.0, 50.0], [12.5, 75.0]]))
    assert rel_err(np.array([a[32, 32], a[48, 8]]), ref).max() <= NN_TOL


def check_cocircular_lattice(lib, O):
    """Regular grid sites: four cocircular sites per cell (degenerate Delaunay, coincident Voronoi vertices), queries on
    grid lines, cell centres (on circumcircles), sites and hull edges."""
    from fastlem_b200 import _native, triangulation
    gx, gy = np.meshgrid(np.arange(30.0), np.arange(20.0))
    sites = np.stack([gx.reshape(-1) * 3.0, gy.reshape(-1) * 5.0], axis=1)
    tri, he = triangulation.delaunay(sites)
    rng = np.random.default_rng(0)
    values = rng.random(sites.shape[0]) * 10.0
    q = np.concatenate([np.stack([rng.random(3000) * 87.0, rng.random(3000) * 95.0], axis=1),
                        np.stack([np.repeat(np.arange(0.0, 87.0, 1.5), 10), np.tile(np.arange(0.0, 95.0, 9.5), 58)], axis=1)])
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        it.set_values(values)
        out = it.points(q)
    ref = O.nn_interpolate(sites, tri, values, q)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ok.sum() > 3000 and rel_err(out[ok], ref[ok]).max() <= NN_TOL


def check_sites_on_a_circle(lib, O):
    """All sites on one circle: every triangle shares the circumcircle, the cavity of the centre is the whole
    triangulation and the interpolant at the centre is the mean.  Beyond FLI_MAX_CAVITY (512) triangles the call fails
    loudly instead of returning a truncated sum."""
    from fastlem_b200 import _native, triangulation
    q = np.array([[50.0, 50.0], [55.0, 48.0]])
    for n in (12, 200):
        th = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
        sites = np.stack([50.0 + 30.0 * np.cos(th), 50.0 + 30.0 * np.sin(th)], axis=1)
        tri, he = triangulation.delaunay(sites)
        values = np.random.default_rng(n).random(n) * 10.0
        with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
            it.set_values(values)
            out = it.points(q)
        assert abs(out[0] - values.mean()) < 1e-9
        assert rel_err(out, O.nn_interpolate(sites, tri, values, q)).max() <= NN_TOL
    th = np.linspace(0.0, 2.0 * np.pi, 700, endpoint=False)
    sites = np.stack([50.0 + 30.0 * np.cos(th), 50.0 + 30.0 * np.sin(th)], axis=1)
    tri, he = triangulation.delaunay(sites)
    with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
        it.set_values(np.zeros(700))
        with pytest.raises(_native.FastlemError) as e:
            it.points(q)
        assert e.value.code == _native.E_INVALID


def check_empty_and_tiny(lib):
    """No triangles at all (0, 1, 2 sites): every query is None; empty query batches and empty row ranges are fine; a
    single triangle reproduces the plane through its vertices (inside, on an edge, on a vertex) and is None outside."""
    from fastlem_b200 import _native
    none = np.zeros(0, dtype=np.uint32)
    for n in (0, 1, 2):
        sites = np.arange(2 * n, dtype=np.float64).reshape(n, 2)
        with _native.Interpolator(sites, none, none, lib_path=lib) as it:
            it.set_values(np.ones(n))
            assert np.isnan(it.points(np.array([[0.5, 0.5], [0.0, 1.0]]))).all()
            assert it.points(np.zeros((0, 2))).shape == (0,)
            assert np.isnan(it.raster(it.raster_desc(3, 2, 0.0, 0.0, 1.0, 1.0))).all()
            assert it.raster(it.raster_desc(3, 2, 0.0, 0.0, 1.0, 1.0, 0.0, 1, 1)).shape == (0, 3)
    sites = np.array([[0.0, 0.0], [4.0, 0.0], [0.0, 4.0]])
    with _native.Interpolator(sites, np.array([0, 1, 2], dtype=np.uint32), np.full(3, 0xFFFFFFFF, dtype=np.uint32),
                              lib_path=lib) as it:
        it.set_values(np.array([1.0, 2.0, 3.0]))  # z = 1 + x/4 + y/2
        out = it.points(np.array([[1.0, 1.0], [2.0, 2.0], [0.0, 0.0], [2.0, 0.0], [5.0, 5.0], [-1.0, 1.0]]))
    assert np.allclose(out[:4], [1.75, 2.5, 1.0, 1.5], rtol=0, atol=1e-12)
    assert np.isnan(out[4:]).all()


def check_large_offsets(lib, O):
    """Sites in large absolute coordinates (map projections): device coordinates are taken relative to the bounding box,
    so the accuracy does not degrade with the offset; values on the sites stay exact."""
    from fastlem_b200 import _native, triangulation
    rng = np.random.default_rng(0)
    base = rng.random((4000, 2)) * 100.0
    values = rng.random(4000) * 100.0
    q0 = 5.0 + rng.random((2000, 2)) * 90.0
    for off in (1.0e3, 5.0e5, 1.0e7):
        sites, q = base + off, q0 + off
        tri, he = triangulation.delaunay(sites)
        with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
            it.set_values(values)
            out = it.points(q)
            assert np.array_equal(it.points(sites[:200]), values[:200])
            img = it.raster(it.raster_desc(16, 16, off, off, 100.0, 100.0, 0.5))
        ref = O.nn_interpolate(sites, tri, values, q)
        assert not np.isnan(ref).any() and not np.isnan(out).any()
        assert rel_err(out, ref).max() <= NN_TOL
        assert np.isfinite(img).mean() > 0.9


def check_fuzz(lib, O, seeds):
    """Random triangulations that stress the geometry: 1000:1 anisotropic boxes, Gaussian clouds, dense clumps inside a
    sparse field, tiny extents at large offsets; random queries, queries next to sites and on edge midpoints."""
    from fastlem_b200 import _native, triangulation
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(4, 500))
        kind, off = seed % 4, [0.0, 1.0e4, -3.0e5, 7.0e6][(seed // 4) % 4]
        if kind == 0:
            sites = rng.random((n, 2)) * [1000.0, 1.0]
        elif kind == 1:
            sites = rng.normal(0.0, 1.0, (n, 2)) * [5.0, 50.0]
        elif kind == 2:
            sites = np.concatenate([rng.random((n, 2)) * 100.0, rng.random((n // 3 + 1, 2)) * 0.01 + 50.0])
        else:
            sites = rng.random((n, 2)) * 1.0e-3
        sites = sites + off
        tri, he = triangulation.delaunay(sites)
        values = rng.random(sites.shape[0]) * 100.0
        lo, hi = sites.min(axis=0), sites.max(axis=0)
        T = tri.reshape(-1, 3)
        k = rng.integers(0, T.shape[0], 50)
        q = np.concatenate([lo + rng.random((300, 2)) * (hi - lo),
                            sites[rng.integers(0, sites.shape[0], 50)] + rng.normal(0.0, 1e-9, (50, 2)) * (hi - lo),
                            0.5 * (sites[T[k, 0]] + sites[T[k, 1]])])
        with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
            it.set_values(values)
            out = it.points(q)
        ref = O.nn_interpolate(sites, tri, values, q)
        # hull-edge midpoints may fall on either side of the hull in double / long double arithmetic
        assert (np.isnan(out) != np.isnan(ref)).sum() <= 3, seed
        ok = ~np.isnan(out) & ~np.isnan(ref)
        # stolen regions bounded by the far-away circumcentres of hull slivers (1000:1 boxes, queries next to a hull
        # site) are ill-conditioned in any formulation: a looser bar than NN_TOL here
        assert ok.sum() > 100 and rel_err(out[ok], ref[ok]).max() <= 1e-6, seed


# ---- CPU tier: emulation build -------------------------------------------------------------------
def test_emu_fuzz(oracle, emu_lib):
    check_fuzz(emu_lib, oracle, range(48))


def test_emu_queries_on_shared_edges(oracle, emu_lib):
    check_queries_on_shared_edges(emu_lib, oracle)


def test_emu_large_offsets(oracle, emu_lib):
    check_large_offsets(emu_lib, oracle)


def test_emu_empty_and_tiny(emu_lib):
    check_empty_and_tiny(emu_lib)


def test_emu_sites_on_a_circle(oracle, emu_lib):
    check_sites_on_a_circle(emu_lib, oracle)


def test_emu_cocircular_lattice(oracle, emu_lib):
    check_cocircular_lattice(emu_lib, oracle)


@pytest.mark.parametrize("path", helpers.nn_golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_emu_matches_golden(emu_lib, path):
    check_golden(emu_lib, path)


@pytest.mark.parametrize("n,seed,bound,lloyd", [(50, 1, (100.0, 100.0), 0), (2000, 2, (100.0, 100.0), 1),
                                                 (5000, 3, (200.0, 100.0), 0), (3, 4, (100.0, 100.0), 0)])
def test_emu_matches_oracle(oracle, emu_lib, n, seed, bound, lloyd):
    check_against_oracle(emu_lib, oracle, n, seed, 3000, bound, lloyd)


def test_emu_properties(emu_lib):
    check_properties(emu_lib)


def test_emu_raster(oracle, emu_lib):
    check_raster(emu_lib, oracle)


def test_emu_orientation_and_validation(emu_lib):
    check_orientation_and_validation(emu_lib)


def test_emu_values_from_solver(oracle, emu_lib):
    check_values_from_solver(emu_lib, oracle)


def test_emu_device_pointer_entry_points(emu_lib):
    """fastlem_interp_set_values_device / fastlem_interp_raster_device: in the emulation build "device" memory is host
    memory, so numpy buffers stand in for the tensors bench.py and ensemble.raster_partitioned pass on the GPU."""
    from fastlem_b200 import _native
    _, sites, tri, he, values = make_case(1200, 81)
    with _native.Interpolator(sites, tri, he, lib_path=emu_lib) as it:
        it.set_values(values)
        desc = it.raster_desc(40, 30, 0.0, 0.0, 100.0, 100.0, 0.5, 5, 25)
        want = it.raster(desc)
        it.set_values(np.zeros_like(values))
        dev_values = np.ascontiguousarray(values)
        it.set_values_device(dev_values.ctypes.data)
        out = np.full((20, 40), -1.0)
        it.raster_device(desc, out.ctypes.data)
    assert np.array_equal(out, want, equal_nan=True)


def test_emu_clustered_sites(oracle, emu_lib):
    """Strongly non-uniform sites: most hint cells are empty (several dilation passes) and walks are long."""
    from fastlem_b200 import _native
    rng = np.random.default_rng(9)
    pts = np.concatenate([rng.normal((20.0, 20.0), 1.0, (1500, 2)), rng.normal((80.0, 70.0), 0.5, (1500, 2)),
                          rng.random((40, 2)) * 100.0])
    m = W.delaunay_model(pts)
    sites, tri, he = W.triangulation_of(m)
    values = np.sin(sites[:, 0] * 0.3) * 10 + sites[:, 1]
    q = queries_for((100.0, 100.0), 3000, 10, margin=-0.05)
    with _native.Interpolator(sites, tri, he, lib_path=emu_lib) as it:
        it.set_values(values)
        out = it.points(q)
        assert it.stats()["grid_passes"] > 1
    ref = oracle.nn_interpolate(sites, tri, values, q)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    # slivers between the clusters have circumradii ~1e3 x the spacing: tolerance relative to the value range
    assert rel_err(out[ok], ref[ok]).max() <= 1e-7


# ---- GPU tier ------------------------------------------------------------------------------------
@pytest.fixture()
def gpu_lib(product_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product has no CPU fallback")
    return product_lib


@pytest.mark.gpu
@pytest.mark.parametrize("path", helpers.nn_golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_gpu_matches_golden(gpu_lib, path):
    check_golden(gpu_lib, path)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,bound,lloyd", [(50, 1, (100.0, 100.0), 0), (2000, 2, (100.0, 100.0), 1),
                                                 (5000, 3, (200.0, 100.0), 0), (3, 4, (100.0, 100.0), 0),
                                                 (20000, 5, (100.0, 100.0), 1)])
def test_gpu_matches_oracle(oracle, gpu_lib, n, seed, bound, lloyd):
    check_against_oracle(gpu_lib, oracle, n, seed, 3000, bound, lloyd)


@pytest.mark.gpu
def test_gpu_queries_on_shared_edges(oracle, gpu_lib):
    check_queries_on_shared_edges(gpu_lib, oracle)


@pytest.mark.gpu
def test_gpu_fuzz(oracle, gpu_lib):
    check_fuzz(gpu_lib, oracle, range(48, 96))


@pytest.mark.gpu
def test_gpu_large_offsets(oracle, gpu_lib):
    check_large_offsets(gpu_lib, oracle)


@pytest.mark.gpu
def test_gpu_empty_and_tiny(gpu_lib):
    check_empty_and_tiny(gpu_lib)


@pytest.mark.gpu
def test_gpu_sites_on_a_circle(oracle, gpu_lib):
    check_sites_on_a_circle(gpu_lib, oracle)


@pytest.mark.gpu
def test_gpu_cocircular_lattice(oracle, gpu_lib):
    check_cocircular_lattice(gpu_lib, oracle)


@pytest.mark.gpu
def test_gpu_properties(gpu_lib):
    check_properties(gpu_lib)


@pytest.mark.gpu
def test_gpu_raster(oracle, gpu_lib):
    check_raster(gpu_lib, oracle)


@pytest.mark.gpu
def test_gpu_orientation_and_validation(gpu_lib):
    check_orientation_and_validation(gpu_lib)


@pytest.mark.gpu
def test_gpu_values_from_solver(oracle, gpu_lib):
    check_values_from_solver(gpu_lib, oracle)


@pytest.mark.gpu
def test_gpu_equals_emulation_bitwise(gpu_lib, emu_lib):
    """Same kernel bodies, no FMA contraction, IEEE division: the device and its host emulation agree bit for bit."""
    from fastlem_b200 import _native
    _, sites, tri, he, values = make_case(4000, 61)
    q = queries_for((100.0, 100.0), 20000, 62)
    outs = []
    for lib in (gpu_lib, emu_lib):
        with _native.Interpolator(sites, tri, he, lib_path=lib) as it:
            it.set_values(values)
            outs.append(it.points(q))
    assert np.array_equal(outs[0], outs[1], equal_nan=True)


@pytest.mark.gpu
def test_gpu_large_raster_properties(gpu_lib):
    """Size-independent checks at a raster size the oracle cannot cover: a plane is reproduced on every pixel inside
    the hull, and row-block partitions reassemble to the full image."""
    from fastlem_b200 import _native
    m = W.delaunay_model(W.random_sites(200000, seed=71))
    sites, tri, he = W.triangulation_of(m)
    with _native.Interpolator(sites, tri, he, lib_path=gpu_lib) as it:
        it.set_values(0.5 * sites[:, 0] + 0.25 * sites[:, 1] - 3.0)
        w = h = 1024
        full = it.raster(it.raster_desc(w, h, 0.0, 0.0, 100.0, 100.0, 0.5))
        cols, rows = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        plane = 0.5 * (100.0 * ((cols + 0.5) / w)) + 0.25 * (100.0 * ((rows + 0.5) / h)) - 3.0
        ok = ~np.isnan(full)
        assert ok.mean() > 0.98
        assert np.abs(full[ok] - plane[ok]).max() < 1e-8
        parts = [it.raster(it.raster_desc(w, h, 0.0, 0.0, 100.0, 100.0, 0.5, r0, r1)) for r0, r1 in ((0, 300), (300, 1024))]
        assert np.array_equal(np.concatenate(parts, axis=0), full, equal_nan=True)
