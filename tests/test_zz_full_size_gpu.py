"""GPU tier, run last (file name): tests added without a GPU at hand in the session that wrote them, and BASELINE
configurations at their full sizes, checked through size-independent
properties because the oracle cannot run them to convergence in test time (tests/helpers.py,
check_converged_properties; the helper itself is validated on the CPU tier in test_emu_parity.py).
C2 at 1M sites lives in test_gpu_parity.py::test_c2_one_million_sites."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_c3_four_million_sites_advanced(oracle, gpu_ctx_factory):
    """BASELINE config C3 at full size: 4M sites, noise-driven erodibility, ocean-mask outlets flood-filled from the rim
    (examples/terrain_generation_advanced.rs:136-210).  The graph is the jittered 2000 x 2000 lattice (a 4M-site
    Delaunay build alone takes minutes on the host; the lattice is the same stand-in DESIGN.md uses for C4).  Every
    stage of iteration 1 against the oracle, then the size-independent properties of the converged result."""
    from tools import workloads as W
    m = W.lattice_model(2000, 2000, jitter=0.35, seed=21)
    p = W.advanced_params(m, seed=3, ocean_level=-0.25)
    outlets = W.outlets_for(m, p)
    assert m["n"] == 4000000 and outlets.size > m["default_outlets"].size, "explicit ocean outlets, not the rim default"
    initial = oracle.initial_elevations(p["base"])
    with gpu_ctx_factory() as ctx:
        _, it = helpers.check_converged_properties(ctx, oracle, m, p, outlets, initial, first_iterations=0)
        assert it > 100


def test_c3_one_million_sites_delaunay_rim_under_ocean_mask(oracle, gpu_ctx_factory):
    """BASELINE config C3 on the real thing at 1M sites: relaxed random sites + the reference's own `add_edge_sites` rim
    (equally spaced boundary sites: every rim edge has the same length, builder.rs:54-131) on a Delaunay graph, outlets =
    the ocean flood-filled from the rim (terrain_generation_advanced.rs:36-42,178-182), noise-driven erodibility.  The
    flood order -- tied rim edges and ~1e5 outlets included -- must come from the device and equal the oracle's heap
    replay; then the size-independent properties of the converged result."""
    from scenarios import scenario
    m, p, outlets, initial, _ = scenario("edge_sites_ocean", 1000000)
    assert outlets.size > 10000
    with gpu_ctx_factory() as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets))
        st = ctx.stats()
        assert st["flood_on_device"] == 1 and st["outlet_ranks_on_device"] == 1
        _, it = helpers.check_converged_properties(ctx, oracle, m, p, outlets, initial, first_iterations=1)
        assert it > 50


def test_c4_sixteen_million_sites(oracle, gpu_ctx_factory):
    """BASELINE config C4 at full size: 16M sites, uniform erodibility, rim outlets, generate() to convergence on one
    B200 (the loop of src/lem/generator.rs:140-210).  The graph is the jittered 4000 x 4000 lattice (planar
    triangulation, degree 6 on average; a 16M-site Delaunay build takes several minutes on the host).  Iteration 1 stage
    by stage against the oracle (receivers, labels, lake connection, areas, response times, elevations: bit-exact),
    then the size-independent properties of the converged result: determinism over two runs, a forest draining to the
    outlets with elevations increasing upstream, conservation of the drainage area, and the oracle's own iteration
    finding the device result to be a fixed point with the same receivers."""
    from tools import workloads as W
    m = W.lattice_model(4000, 4000, jitter=0.35, seed=1)
    p = W.uniform_params(m["n"])
    outlets = W.outlets_for(m, p)
    assert m["n"] == 16000000
    initial = oracle.initial_elevations(p["base"])
    with gpu_ctx_factory() as ctx:
        e, it = helpers.check_converged_properties(ctx, oracle, m, p, outlets, initial, first_iterations=0)
        st = ctx.stats()
        assert it > 1000
        assert st["flood_on_device"] == 1
        # the lock-free hand-offs of K4 / K5 under load: a third run of the whole thing, identical bits
        e3, it3 = ctx.generate()
        assert it3 == it and np.array_equal(e3, e)


def test_stress_two_hundred_generates_one_million_sites(product_lib):
    """The sweeps of K4 and K5 hand work between warps without locks (release / acquire through L2, fl_flow.cuh,
    fl_elev.cuh): 200 back-to-back generate() of the C2 terrain (1M sites, about a thousand iterations each, 2e5
    launches of the dataflow kernels) must give identical bits and identical iteration counts every time."""
    import hashlib
    from fastlem_b200 import _native
    from scenarios import scenario
    m, p, outlets, initial, _ = scenario("uniform", 1000000)
    with _native.Context(0, product_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        seen = set()
        for _ in range(200):
            e, it = ctx.generate()
            seen.add((it, hashlib.sha1(e.tobytes()).hexdigest()))
        assert len(seen) == 1, f"{len(seen)} different results in 200 runs"


def test_ensemble_members_on_one_shared_graph_gpu(oracle, product_lib):
    """The shared-graph ensemble runner on the device (same checks as tests/test_emu_parity.py)."""
    import numpy as np
    from fastlem_b200 import _native, ensemble
    from scenarios import scenario
    m, p, outlets, initial, _ = scenario("uniform", 30000)
    n = m["n"]
    rng = np.random.default_rng(4)
    other_outlets = np.unique(np.concatenate([outlets[::2], rng.integers(0, n, 5).astype(np.uint32)])).astype(np.uint32)
    members = [
        dict(initial=initial, erodibility=p["erodibility"], uplift=p["uplift"], outlets=outlets),
        dict(initial=initial, erodibility=0.5 + rng.random(n), uplift=p["uplift"], outlets=outlets),
        dict(initial=initial, erodibility=0.5 + rng.random(n), uplift=p["uplift"], outlets=other_outlets),
        dict(initial=initial, erodibility=p["erodibility"], uplift=p["uplift"], outlets=outlets,
             tan_max_slope=helpers.tan_of(np.full(n, 0.3))),
    ]
    res = ensemble.run_members_shared_graph(m, members, lambda: _native.Context(0, product_lib), max_iteration=60)
    for k, (mem, (e, it)) in enumerate(zip(members, res)):
        ms = None if "tan_max_slope" not in mem else np.full(n, 0.3)
        ref, ref_it = oracle.generate(m, mem["erodibility"], mem["uplift"], ms, mem["outlets"], mem["initial"], 60)
        assert it == ref_it, k
        assert np.array_equal(e, ref), f"member {k}"


def test_random_graphs_bit_exact_gpu(oracle, product_lib):
    """The differential fuzz of tests/test_fuzz.py on the device: random small graphs (ties, several components, random
    outlets, mixed max_slope), every sweep implementation, bit-exact against the oracle."""
    import numpy as np
    from fastlem_b200 import _native
    from tools.fuzz_solver import random_case
    done = 0
    for seed in range(200, 280):
        case = random_case(seed, oracle)
        if case is None:
            continue
        m, p, outlets, initial, max_iteration = case
        ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, max_iteration)
        for sweep in (0, 1, 2, 3):
            with _native.Context(0, product_lib) as ctx:
                ctx.set_option("sweep", sweep)
                helpers.load_ctx(ctx, m, p, outlets, initial)
                e, it = ctx.generate(max_iteration)
            assert it == ref_it, (seed, sweep)
            assert np.array_equal(e, ref, equal_nan=True), (seed, sweep)
        done += 1
    assert done >= 40


def test_context_reuse_sequences_gpu(oracle, product_lib):
    """tests/test_fuzz.py::test_context_reuse_sequences on the device (other seeds)."""
    import numpy as np
    from fastlem_b200 import _native
    from tools.fuzz_solver import random_graph
    for seed in range(100, 120):
        rng = np.random.default_rng(seed)
        graphs = [random_graph(rng, 5, 200) for _ in range(2)]
        with _native.Context(0, product_lib) as ctx:
            m = outlets = None
            for step in range(8):
                if step == 0 or rng.random() < 0.25:
                    m = graphs[int(rng.integers(0, 2))]
                    n = m["n"]
                    ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
                    outlets = None
                if outlets is None or rng.random() < 0.4:
                    outlets = np.sort(rng.choice(n, int(rng.integers(1, max(2, n // 6))), replace=False)).astype(np.uint32)
                ms = None
                if rng.random() < 0.4:
                    ms = 0.05 + rng.random(n) * 0.8
                    ms[rng.random(n) < 0.3] = np.nan
                k = 0.2 + rng.random(n) * 2
                u = np.ones(n) if rng.random() < 0.5 else 0.5 + rng.random(n)
                initial = oracle.initial_elevations(np.zeros(n) if rng.random() < 0.7 else rng.random(n))
                mi = int(rng.integers(1, 50))
                if rng.random() < 0.3:
                    ctx.set_option("sweep", int(rng.integers(0, 4)))
                ctx.set_parameters(initial, k, u, helpers.tan_of(ms), outlets)
                e, it = ctx.generate(mi)
                ref, ref_it = oracle.generate(m, k, u, ms, outlets, initial, mi)
                assert it == ref_it and np.array_equal(e, ref, equal_nan=True), (seed, step)


def test_special_parameter_values_gpu(oracle, product_lib):
    """tests/test_fuzz.py::check_special_parameter_values on the device (inf / NaN propagation, other seeds)."""
    from test_fuzz import check_special_parameter_values
    check_special_parameter_values(product_lib, oracle, range(60, 120))
