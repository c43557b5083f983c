"""GPU tier (-m gpu): the parity gate.  Everything goes through the C ABI of the nvcc-built product library
(fastlem_b200/_lib/libfastlem_b200.so) and is compared with the CPU oracle on identical inputs.

Bars (north_star): receivers, basin labels, lake connection and traversal membership bit-exact; drainage
areas, response times and elevations within 1e-9 relative (helpers.REL_TOL) -- in practice the CUDA path
reproduces them bit for bit, which the tests also record.
"""
import numpy as np
import pytest

import helpers
from scenarios import SMALL, scenario

pytestmark = pytest.mark.gpu


def test_library_is_the_cuda_build(gpu_ctx_factory):
    with gpu_ctx_factory() as ctx:
        assert ctx.version().endswith("sm_100a")


@pytest.mark.parametrize("name", SMALL)
def test_first_iteration_stages(oracle, gpu_ctx_factory, name):
    m, p, outlets, initial, _ = scenario(name)
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        _, exact = helpers.check_first_iteration(ctx, oracle, m, p, outlets, initial)
        assert exact, "f64 stages are within tolerance but not bit-identical"


@pytest.mark.parametrize("sweep", [0, 1, 2, 3])
@pytest.mark.parametrize("name", SMALL)
def test_generate_to_convergence(oracle, gpu_ctx_factory, name, sweep):
    m, p, outlets, initial, max_iteration = scenario(name)
    with gpu_ctx_factory(sweep=sweep) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("sweep", [0, 1, 2, 3])
@pytest.mark.parametrize("path", helpers.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_vectors(gpu_ctx_factory, path, sweep):
    with gpu_ctx_factory(sweep=sweep) as ctx:
        helpers.check_against_golden(ctx, path)


@pytest.mark.parametrize("every", [0, 1, 3, 1000])
@pytest.mark.parametrize("name", ["uniform", "max_slope", "uplift", "advanced", "disconnected", "lattice_regular"])
def test_dataflow_sweeps_on_stale_numbering(oracle, gpu_ctx_factory, name, every):
    """sweep 3 keeps a site numbering for several iterations; segments are whatever chains are still contiguous.
    rebuild_every: 0 adaptive, 1 every iteration, 3 every third, 1000 never after the first."""
    m, p, outlets, initial, max_iteration = scenario(name)
    with gpu_ctx_factory(sweep=3, rebuild_every=every) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("park_after", [0, 8, 16, 64])
@pytest.mark.parametrize("name", ["uniform", "max_slope", "uplift"])
def test_dataflow_long_chain_hand_over(oracle, gpu_ctx_factory, name, park_after):
    """Long chains are parked by the thread-level pass and continued by warps (park_after = 0: never)."""
    m, p, outlets, initial, max_iteration = scenario(name, 30000)
    with gpu_ctx_factory(sweep=3, park_after=park_after) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("n", [30000, 200000])
def test_dataflow_sweeps_repeatable_at_size(oracle, gpu_ctx_factory, n):
    """The dataflow kernel's hand-offs race differently from run to run; results must not."""
    m, p, outlets, initial, _ = scenario("uniform", n)
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], None, outlets, initial)
    with gpu_ctx_factory(sweep=3) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        for _ in range(3):
            e, it = ctx.generate()
            assert it == ref_it and np.array_equal(e, ref)


@pytest.mark.parametrize("k", [0, 1, 2, 5])
def test_max_iteration(oracle, gpu_ctx_factory, k):
    m, p, outlets, initial, _ = scenario("uniform")
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        e, it = ctx.generate(k)
        ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, k)
        assert it == ref_it == k
        assert np.array_equal(e, ref)


@pytest.mark.parametrize("sweep", [1, 2, 3])
@pytest.mark.parametrize("name", ["uniform", "max_slope", "uplift", "disconnected", "hub", "interior_outlets"])
def test_stages_after_several_iterations_on_paths(oracle, gpu_ctx_factory, name, sweep):
    """Stage dumps in the path layout (renumbered sites) map back to the caller's numbering."""
    m, p, outlets, initial, _ = scenario(name)
    k = 4
    e = initial.copy()
    for _ in range(k - 1):
        e = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)["elevations"]
    ref = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)
    with gpu_ctx_factory(sweep=sweep, keep_stages=1) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        out, it = ctx.generate(k)
        assert it == k
        assert np.array_equal(out, ref["elevations"], equal_nan=True)
        assert np.array_equal(ctx.fetch("receivers"), ref["next"])
        assert np.array_equal(ctx.fetch("receivers_initial"), ref["next_initial"])
        assert np.array_equal(ctx.fetch("labels_initial"), ref["subroot"])
        assert np.array_equal(ctx.fetch("depth") != 0xFFFFFFFF, ref["order"] != oracle.NONE)
        assert np.array_equal(ctx.fetch("drainage_area"), ref["drainage"])
        assert np.array_equal(ctx.fetch("response_time"), ref["response"])


def test_c1_landscape_evolution_as_shipped(oracle, gpu_ctx_factory):
    """BASELINE config C1: examples/landscape_evolution.rs (30 000 sites, relaxate_sites(1), k = 1, hull outlets)."""
    m, p, outlets, initial, _ = scenario("rust_sites", 30000)
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        helpers.check_first_iteration(ctx, oracle, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, None)


def test_max_slope_100k_long_paths(oracle, gpu_ctx_factory):
    """The clamp chain of the warp-per-path kernel on long paths (tests/landscape_evolution.rs scenario, larger)."""
    m, p, outlets, initial, _ = scenario("max_slope", 100000)
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, None)


def test_c3_style_advanced_200k(oracle, gpu_ctx_factory):
    """terrain_generation_advanced-style (noise erodibility + ocean-mask outlets) at 200k sites, to convergence."""
    m, p, outlets, initial, _ = scenario("advanced", 200000)
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, None)


def test_c7_nonuniform_uplift_lakes_every_iteration(oracle, gpu_ctx_factory):
    m, p, outlets, initial, _ = scenario("uplift", 100000)
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, 50)
        assert ctx.stats()["lake_iterations"] > 1


@pytest.mark.parametrize("opts", [dict(k5_split=0, key_base=1), dict(k5_split=0, fuse_levels=0), dict(incremental=0), dict(incr_div=1),
                                  dict(flood_device=0), dict(first_flow=0), dict(k5_split=0), dict(k5_cut=0), dict(k5_cut=1), dict(k5_cut=3), dict(k5_cut=8), dict(k5_top_cap=50),
                                  dict(k1_bulk=0), dict(overlap=0), dict(outlet_closed_form=0), dict(k1_bulk=0, overlap=0, k5_cut=2),
                                  dict(rebuild_growth=1, rebuild_height=100), dict(rebuild_growth=50, rebuild_height=400)])
@pytest.mark.parametrize("name,n", [("uniform", 30000), ("advanced", 20000), ("max_slope", 20000)])
def test_solver_options_do_not_change_results(oracle, gpu_ctx_factory, name, n, opts):
    """Every scheduling option (deep-nesting ordering path, per-level K5 launches, full K4 every iteration, incremental
    K4 whenever possible, host flood replay) must give the oracle's bits."""
    m, p, outlets, initial, max_iteration = scenario(name, n)
    with gpu_ctx_factory(**opts) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("name,n", [("uniform", 3000), ("advanced", 5000), ("lattice", None), ("lattice_regular", None),
                                    ("disconnected", None), ("interior_outlets", 8000), ("single_outlet", 6000),
                                    ("uniform", 200000), ("advanced", 200000), ("uniform", 1000000),
                                    ("edge_sites_ocean", 3000), ("edge_sites_ocean", 100000), ("edge_sites_partial", 3000)])
def test_flood_order_on_device(oracle, gpu_ctx_factory, name, n):
    """fl_floodgpu.cuh: pop order of the lake flood from the minimum spanning tree, against the oracle's heap replay."""
    m, p, outlets, initial, _ = scenario(name, n) if n else scenario(name)
    ref = oracle.flood_order(m, outlets)
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), ref)
        st = ctx.stats()
        assert st["flood_on_device"] == (0 if name in ("lattice_regular", "edge_sites_partial") else 1)
        assert st["outlet_ranks_on_device"] == st["flood_on_device"]  # closed form of the outlets' own ranks
    with gpu_ctx_factory(flood_device=0) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), ref)
        assert ctx.stats()["flood_on_device"] == 0


def test_c2_one_million_sites(oracle, gpu_ctx_factory):
    """BASELINE config C2 (1M random sites, uniform erodibility, hull outlets): the oracle is too slow to
    converge here (~10 min), so compare the first iterations against it, then check size-independent
    properties of the converged GPU result."""
    m, p, outlets, initial, _ = scenario("uniform", 1000000)
    with gpu_ctx_factory() as ctx:
        helpers.check_converged_properties(ctx, oracle, m, p, outlets, initial)


def test_host_mirror_on_gpu(oracle, product_lib):
    import fastlem_b200 as fl
    m, p, outlets, initial, _ = scenario("max_slope")
    params = [fl.TopographicalParameters.default().set_max_slope(3.14 * 0.1) for _ in range(m["n"])]
    gen = fl.TerrainGenerator.default().set_model(fl.TerrainModel2D.from_workload(m)).set_parameters(params)
    terrain = gen.generate()
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial)
    assert gen.last_iterations == ref_it
    assert np.array_equal(terrain.elevations(), ref)


def test_two_contexts_are_independent(oracle, gpu_ctx_factory):
    a = scenario("uniform")
    b = scenario("advanced")
    with gpu_ctx_factory() as ca, gpu_ctx_factory() as cb:
        helpers.load_ctx(ca, a[0], a[1], a[2], a[3])
        helpers.load_ctx(cb, b[0], b[1], b[2], b[3])
        ea, _ = ca.generate()
        eb, _ = cb.generate()
        ea2, _ = ca.generate()
    assert np.array_equal(ea, ea2)
    assert np.array_equal(ea, oracle.generate(a[0], a[1]["erodibility"], a[1]["uplift"], None, a[2], a[3])[0])
    assert np.array_equal(eb, oracle.generate(b[0], b[1]["erodibility"], b[1]["uplift"], None, b[2], b[3])[0])
