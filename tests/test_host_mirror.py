"""CPU tier: the Python mirror of the reference's public interface for this path
(TerrainGenerator / TopographicalParameters / TerrainModel2D / Terrain2D / GenerationError), run on the
host emulation build.  The scenarios follow the reference's own tests/ and examples/."""
import numpy as np
import pytest

import fastlem_b200 as fl
from scenarios import scenario


def _generator(emu_lib):
    g = fl.TerrainGenerator.default()
    g._lib_path = emu_lib
    return g


def test_errors_match_generation_error_variants(emu_lib):
    m, *_ = scenario("uniform", 800)
    model = fl.TerrainModel2D.from_workload(m)
    with pytest.raises(fl.ModelNotSet):
        _generator(emu_lib).generate()
    with pytest.raises(fl.ParametersNotSet):
        _generator(emu_lib).set_model(model).generate()
    with pytest.raises(fl.InvalidNumberOfParameters):
        _generator(emu_lib).set_model(model).set_parameters([fl.TopographicalParameters.default()] * 3).generate()
    assert issubclass(fl.ModelNotSet, fl.GenerationError)


def test_landscape_evolution_example(emu_lib, oracle):
    """examples/landscape_evolution.rs:18-34 with fewer sites."""
    m, p, outlets, initial, _ = scenario("uniform", 800)
    num = m["n"]
    terrain = _generator(emu_lib).set_model(fl.TerrainModel2D.from_workload(m)).set_parameters(
        [fl.TopographicalParameters.default().set_erodibility(1.0) for _ in range(num)]).generate()
    ref, _ = oracle.generate(m, p["erodibility"], p["uplift"], None, outlets, initial)
    assert np.array_equal(terrain.elevations(), ref)
    assert terrain.sites().shape == (num, 2)


def test_landscape_evolution_test_scenario(emu_lib, oracle):
    """tests/landscape_evolution.rs:19-32: every setter, max_slope = Some(3.14 * 0.1)."""
    m, p, outlets, initial, _ = scenario("max_slope", 900)
    params = [fl.TopographicalParameters.default().set_base_elevation(0.0).set_erodibility(1.0).set_uplift_rate(1.0)
              .set_is_outlet(False).set_max_slope(3.14 * 0.1) for _ in range(m["n"])]
    gen = _generator(emu_lib).set_model(fl.TerrainModel2D.from_workload(m)).set_parameters(params)
    terrain = gen.generate()
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial)
    assert gen.last_iterations == ref_it
    assert np.array_equal(terrain.elevations(), ref)


def test_tan_of_max_slope_is_libm(emu_lib):
    """generator.rs:194 `max_slope.tan()`: the mirror's tan is libm's (math.tan), bit for bit, NaN (None) kept."""
    import math
    from fastlem_b200 import _native
    rng = np.random.default_rng(3)
    ms = rng.uniform(-1.6, 1.6, 200000)
    ms[::97] = np.nan
    got = _native.host_tan_max_slope(ms, emu_lib)
    want = np.array([math.tan(v) if v == v else np.nan for v in ms])
    assert np.array_equal(got, want, equal_nan=True)


def test_random_per_site_max_slope_is_bit_exact(emu_lib, oracle):
    """Per-site Some(max_slope) with random angles (round-1 advisor finding: np.tan differs from libm's tan by an ulp
    for ~0.5 % of the inputs, and one ulp in tan breaks bit-exactness of every clamped site)."""
    m, p, outlets, initial, _ = scenario("uniform", 1500)
    rng = np.random.default_rng(11)
    ms = rng.uniform(0.02, 1.2, m["n"])
    ms[rng.random(m["n"]) < 0.2] = np.nan  # None
    arrays = fl.ParameterArrays(p["base"], p["erodibility"], p["uplift"], np.zeros(m["n"], dtype=bool), max_slope=ms)
    gen = _generator(emu_lib).set_model(fl.TerrainModel2D.from_workload(m)).set_parameters(arrays).set_max_iteration(40)
    terrain = gen.generate()
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], ms, outlets, initial, 40)
    assert gen.last_iterations == ref_it
    assert np.array_equal(terrain.elevations(), ref)


def test_explicit_outlets_override_default(emu_lib, oracle):
    m, p, outlets, initial, _ = scenario("interior_outlets")
    arrays = fl.ParameterArrays(p["base"], p["erodibility"], p["uplift"], p["is_outlet"])
    terrain = _generator(emu_lib).set_model(fl.TerrainModel2D.from_workload(m)).set_parameters(arrays) \
        .set_max_iteration(7).generate()
    ref, _ = oracle.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, 7)
    assert np.array_equal(terrain.elevations(), ref)


def test_noise_is_the_reference_stream(emu_lib, oracle):
    from fastlem_b200 import _native
    base = np.linspace(-1.0, 1.0, 257)
    assert np.array_equal(_native.host_initial_elevations(base, emu_lib), oracle.initial_elevations(base))


def test_get_elevation_render_loop(emu_lib, oracle):
    """examples/landscape_evolution.rs:36-62: generate(), then get_elevation per pixel (`if let Some(e)`), against the
    oracle's interpolation of the oracle's elevations; the one-call raster gives the same image."""
    from tools import workloads as W
    m = W.delaunay_model(W.random_sites(700, seed=5), lloyd=1, bound_min=(0, 0), bound_max=(100, 100))
    model = fl.TerrainModel2D.from_workload(m)
    num = model.num()
    terrain = _generator(emu_lib).set_model(model).set_parameters(
        [fl.TopographicalParameters.default() for _ in range(num)]).generate()
    bound_max = fl.Site2D(100.0, 100.0)
    img_width = img_height = 24
    image = np.full((img_height, img_width), np.nan)
    for imgx in range(img_width):
        for imgy in range(img_height):
            x = bound_max.x * (imgx / img_width)
            y = bound_max.y * (imgy / img_height)
            e = terrain.get_elevation(fl.Site2D(x, y))
            if e is not None:
                image[imgy, imgx] = e
    raster = terrain.raster(img_width, img_height, 0.0, 0.0, bound_max.x, bound_max.y)
    assert np.array_equal(image, raster, equal_nan=True)
    sites, tri, _ = W.triangulation_of(m)
    cols, rows = np.meshgrid(np.arange(img_width), np.arange(img_height))
    q = np.stack([100.0 * (cols.reshape(-1) / img_width), 100.0 * (rows.reshape(-1) / img_height)], axis=1)
    ref = oracle.nn_interpolate(sites, tri, terrain.elevations(), q).reshape(img_height, img_width)
    assert np.array_equal(np.isnan(ref), np.isnan(raster))
    ok = ~np.isnan(ref)
    assert ok.sum() > 400 and np.isnan(ref[0, 0])  # the corner pixel (0, 0) lies outside the hull of random sites
    assert (np.abs(raster[ok] - ref[ok]) <= 1e-9 * np.maximum(1.0, np.abs(ref[ok]))).all()
    assert terrain.get_elevation(fl.Site2D(-1.0, 50.0)) is None


def test_interpolator_triangulates_on_demand(emu_lib, oracle):
    """TerrainInterpolator2D::new(sites) without a builder triangulation: a host Delaunay of the sites on first use."""
    from fastlem_b200 import triangulation
    rng = np.random.default_rng(3)
    sites = rng.random((300, 2)) * 50.0
    values = sites[:, 0] * 0.1 + np.sin(sites[:, 1])
    it = fl.TerrainInterpolator2D(sites, lib_path=emu_lib)
    assert it.stats() is None  # nothing built yet
    q = 10.0 + rng.random((50, 2)) * 30.0
    out = it.interpolate_many(values, q)
    tri, he = triangulation.delaunay(sites)
    ref = oracle.nn_interpolate(sites, tri, values, q)
    assert (np.abs(out - ref) <= 1e-9 * np.maximum(1.0, np.abs(ref))).all()
    assert it.interpolate(values, fl.Site2D(q[0, 0], q[0, 1])) == out[0]
    it.close()


def test_advanced_example_render(emu_lib, oracle):
    """examples/terrain_generation_advanced.rs:285-315 (two displaced get_elevation calls per pixel) through
    examples/terrain_generation_advanced.py, against per-pixel oracle queries at the example's coordinates."""
    import importlib.util
    import math
    import os
    from scenarios import ROOT
    from tools import workloads as W
    spec = importlib.util.spec_from_file_location("adv", os.path.join(ROOT, "examples", "terrain_generation_advanced.py"))
    adv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(adv)
    size = 20
    terrain, elevation, brightness = adv.main(1500, size, None, lib_path=emu_lib)
    m = W.delaunay_model(W.random_sites(1500, (0.0, 0.0), (100.0, 100.0), seed=0), lloyd=1, bound_min=(0.0, 0.0),
                         bound_max=(100.0, 100.0))
    sites, tri, _ = W.triangulation_of(m)
    sdx, sdy = 0.3 * math.cos(3.14 * 0.25), 0.3 * math.sin(3.14 * 0.25)
    for imgx, imgy in ((3, 4), (10, 10), (19, 0), (0, 19), (7, 15)):
        x = (100.0 - sdx) * ((imgx + 0.5) / size) + 0.0
        y = (100.0 - sdy) * ((imgy + 0.5) / size) + 0.0
        e1, e2 = oracle.nn_interpolate(sites, tri, terrain.elevations(), np.array([[x, y], [x + sdx, y + sdy]]))
        if np.isnan(e1) or np.isnan(e2):
            assert np.isnan(elevation[imgy, imgx]) and np.isnan(brightness[imgy, imgx])
            continue
        assert abs(elevation[imgy, imgx] - e1) <= 1e-9 * max(1.0, abs(e1))
        want = 1.0 - math.sin(math.atan((e1 - e2) / 50.0))
        assert abs(brightness[imgy, imgx] - want) <= 1e-9
    assert np.isfinite(elevation).mean() > 0.8


def test_interpolator_reads_the_slice_on_every_call(emu_lib):
    """interpolator.rs:17-27 reads `elevations` on every call: values changed in place must be seen (the upload is only
    reused for arrays that cannot change), and a NaN elevation gives Some(NaN), not the None of a query outside the hull."""
    rng = np.random.default_rng(5)
    sites = rng.random((200, 2)) * 50.0
    values = np.full(200, 2.0)
    it = fl.TerrainInterpolator2D(sites, lib_path=emu_lib)
    inside, outside = fl.Site2D(25.0, 25.0), fl.Site2D(-5.0, 25.0)
    assert abs(it.interpolate(values, inside) - 2.0) < 1e-12
    values[:] = 7.0  # same object, new contents
    assert abs(it.interpolate(values, inside) - 7.0) < 1e-12
    assert it.interpolate(values, outside) is None
    values[:] = np.nan
    z = it.interpolate(values, inside)
    assert z is not None and z != z  # Some(NaN)
    assert it.interpolate(values, outside) is None
    frozen = np.full(200, 3.0)
    frozen.setflags(write=False)
    assert abs(it.interpolate(frozen, inside) - 3.0) < 1e-12
    assert abs(it.interpolate(frozen, inside) - 3.0) < 1e-12  # (served from the uploaded copy)
    it.close()
