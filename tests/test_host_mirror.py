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
