"""CPU tier: the solver's host orchestration and kernel bodies, compiled serially for the host (-DFL_EMU),
against the oracle.  This checks LOGIC before GPU time is spent; the GPU tier (test_gpu_parity.py) is the
parity gate proper."""
import numpy as np
import pytest

import helpers
from scenarios import SMALL, scenario


def _ctx(emu_lib, **opts):
    from fastlem_b200 import _native
    ctx = _native.Context(0, emu_lib)
    assert ctx.version().endswith("emu")
    for k, v in opts.items():
        ctx.set_option(k, v)
    return ctx


SWEEPS = [0, 1, 2, 3]


@pytest.mark.parametrize("name", SMALL)
def test_first_iteration_stages(oracle, emu_lib, name):
    m, p, outlets, initial, _ = scenario(name)
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        _, exact = helpers.check_first_iteration(ctx, oracle, m, p, outlets, initial)
        assert exact


@pytest.mark.parametrize("sweep", SWEEPS)
@pytest.mark.parametrize("name", SMALL)
def test_generate(oracle, emu_lib, name, sweep):
    m, p, outlets, initial, max_iteration = scenario(name)
    with _ctx(emu_lib, sweep=sweep) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)
        if sweep == 1 and ctx.stats()["iterations"] > 1:
            assert ctx.stats()["rebuilds"] >= 1


@pytest.mark.parametrize("sweep", [1, 3])
@pytest.mark.parametrize("name", ["uniform", "max_slope", "uplift", "disconnected", "hub", "interior_outlets"])
def test_stages_after_several_iterations_on_paths(oracle, emu_lib, name, sweep):
    """Stage dumps in the path layout (renumbered sites) map back to the caller's numbering."""
    m, p, outlets, initial, _ = scenario(name)
    k = 4
    e = initial.copy()
    for _ in range(k - 1):
        e = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)["elevations"]
    ref = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)
    with _ctx(emu_lib, sweep=sweep, keep_stages=1) as ctx:
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


@pytest.mark.parametrize("sweep", SWEEPS)
@pytest.mark.parametrize("path", helpers.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden(emu_lib, path, sweep):
    with _ctx(emu_lib, sweep=sweep) as ctx:
        helpers.check_against_golden(ctx, path)


@pytest.mark.parametrize("every", [0, 1, 3, 1000])
@pytest.mark.parametrize("name", ["uniform", "max_slope", "uplift", "advanced", "disconnected", "lattice_regular"])
def test_dataflow_sweeps_on_stale_numbering(oracle, emu_lib, name, every):
    """sweep 3 keeps a site numbering for several iterations; segments are whatever chains are still contiguous.
    rebuild_every: 0 adaptive, 1 every iteration, 3 every third, 1000 never after the first."""
    m, p, outlets, initial, max_iteration = scenario(name)
    with _ctx(emu_lib, sweep=3, rebuild_every=every) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("park_after", [0, 8, 64])
def test_dataflow_parking(oracle, emu_lib, park_after):
    m, p, outlets, initial, max_iteration = scenario("uniform")
    with _ctx(emu_lib, sweep=3, park_after=park_after) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("k", [0, 1, 3])
def test_max_iteration(oracle, emu_lib, k):
    m, p, outlets, initial, _ = scenario("uniform")
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        e, it = ctx.generate(k)
        ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, k)
        assert it == ref_it == k
        assert np.array_equal(e, ref)
        if k == 0:
            assert np.array_equal(e, initial)  # generator.rs:140: zero bodies -> the noise comes back


@pytest.mark.parametrize("device", [0, 1])
@pytest.mark.parametrize("name", SMALL)
def test_flood_rank_matches_oracle(oracle, emu_lib, name, device):
    """Pop order of the lake flood (stream_tree.rs:175-243): exact host replay (flood_device=0) and the spanning-tree
    formulation on the device (fl_floodgpu.cuh), which must fall back to the replay when edge lengths tie."""
    m, p, outlets, initial, _ = scenario(name)
    with _ctx(emu_lib, flood_device=device) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets)), name
        on_device = ctx.stats()["flood_on_device"]
        assert on_device == (1 if device and name not in ("lattice_regular",) else 0)
        assert ctx.stats()["outlet_ranks_on_device"] == on_device  # closed form of the outlets' own ranks
    if device:  # the same with the outlets' prefix replayed on the host
        with _ctx(emu_lib, outlet_closed_form=0) as ctx:
            helpers.load_ctx(ctx, m, p, outlets, initial)
            assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets)), name
            assert ctx.stats()["outlet_ranks_on_device"] == 0


@pytest.mark.parametrize("count", list(range(1, 41)) + [63, 64, 65, 127, 128, 129, 1000])
def test_outlet_ranks_closed_form_every_count(oracle, emu_lib, count):
    """The outlets' own ranks (BinaryHeap at equal keys 0.0, stream_tree.rs:184-197) from their closed form on the device
    (fl_floodgpu.cuh, k_flg_outlet_*), for every small number of outlets (every shape of the heap's last level) and a
    few larger ones, outlets in a shuffled order."""
    m, p, _, initial, _ = scenario("uniform", 2500)
    rng = np.random.default_rng(count)
    outlets = rng.choice(m["n"], count, replace=False).astype(np.uint32)  # unsorted: user-supplied default_outlets order
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets))
        st = ctx.stats()
        assert st["flood_on_device"] == 1 and st["outlet_ranks_on_device"] == 1


def test_outlet_ranks_closed_form_refused(oracle, emu_lib):
    """An outlet without neighbours at the front of the order: its pop pushes nothing, the next pop moves a 0.0 entry from
    the tail and the closed form does not hold -- the validity check must send the prefix to the host replay."""
    m, p, outlets, initial, _ = scenario("uniform", 1500)
    n = m["n"] + 3  # three sites without any edge, placed at the front of the outlet list
    m2 = dict(m, n=n, row_ptr=np.concatenate([m["row_ptr"], np.full(3, m["row_ptr"][-1], np.uint32)]).astype(np.uint32),
              areas=np.concatenate([m["areas"], np.ones(3)]))
    p2 = dict(erodibility=np.ones(n), uplift=np.ones(n), max_slope=None)
    outlets2 = np.concatenate([np.arange(n - 3, n, dtype=np.uint32), outlets]).astype(np.uint32)
    initial2 = oracle.initial_elevations(np.zeros(n))
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m2, p2, outlets2, initial2)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m2, outlets2))
        st = ctx.stats()
        assert st["flood_on_device"] == 1 and st["outlet_ranks_on_device"] == 0
        assert helpers.check_generate(ctx, oracle, m2, p2, outlets2, initial2, 30)


@pytest.mark.parametrize("name,n", [("uniform", 20000), ("advanced", 30000), ("interior_outlets", 8000),
                                    ("single_outlet", 6000)])
def test_flood_rank_device_larger(oracle, emu_lib, name, n):
    m, p, outlets, initial, _ = scenario(name, n)
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets))
        assert ctx.stats()["flood_on_device"] == 1


@pytest.mark.parametrize("name,on_device", [("edge_sites_ocean", 1), ("edge_sites_partial", 0)])
def test_flood_rank_with_add_edge_sites_rim(oracle, emu_lib, name, on_device):
    """The reference's own `add_edge_sites` rim (builder.rs:54-131: equally spaced boundary sites, exact edge-length ties)
    under an ocean mask flood-filled from the rim (terrain_generation_advanced.rs:178-182).  When the whole rim is ocean
    every tied edge joins two outlets -- a self-loop of the contracted source that only ever yields stale heap entries --
    and the flood order is computed on the device; a rim that is partly land keeps real ties and takes the host replay.
    Either way the order and the whole generate() equal the oracle's."""
    m, p, outlets, initial, max_iteration = scenario(name)
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets))
        assert ctx.stats()["flood_on_device"] == on_device
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


def test_rerun_restarts_from_initial(emu_lib):
    m, p, outlets, initial, _ = scenario("uniform", 800)
    with _ctx(emu_lib) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        a, ita = ctx.generate()
        b, itb = ctx.generate()
        assert ita == itb and np.array_equal(a, b)


@pytest.mark.parametrize("incr_div", [1, 4, 16])
@pytest.mark.parametrize("every", [0, 3, 1000])
@pytest.mark.parametrize("name", ["uniform", "max_slope", "uplift", "advanced", "disconnected", "lattice_regular",
                                  "interior_outlets", "single_outlet", "base_field"])
def test_incremental_area_update(oracle, emu_lib, name, every, incr_div):
    """K4 redone only above re-routed sites (fl_flow.cuh, 'Incremental K4'): same bits, same iteration count.
    incr_div = 1 takes the incremental pass whenever the previous state is reusable."""
    m, p, outlets, initial, max_iteration = scenario(name)
    with _ctx(emu_lib, sweep=3, rebuild_every=every, incr_div=incr_div) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)
        st = ctx.stats()
        if name in ("uniform", "advanced", "max_slope") and every != 1:
            assert st["incremental_iterations"] > 0
    with _ctx(emu_lib, sweep=3, rebuild_every=every, incremental=0) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)
        assert ctx.stats()["incremental_iterations"] == 0


def test_incremental_stages_midway(oracle, emu_lib):
    """Stage dumps (areas, response times, receivers) after an incremental iteration."""
    m, p, outlets, initial, _ = scenario("uniform", 6000)
    for k in (6, 9, 15):
        e = initial.copy()
        for _ in range(k - 1):
            e = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)["elevations"]
        ref = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)
        with _ctx(emu_lib, sweep=3, incr_div=1, rebuild_every=1000) as ctx:
            helpers.load_ctx(ctx, m, p, outlets, initial)
            out, it = ctx.generate(k)
            assert it == k and ctx.stats()["incremental_iterations"] >= k - 3
            assert np.array_equal(out, ref["elevations"])
            assert np.array_equal(ctx.fetch("receivers"), ref["next"])
            assert np.array_equal(ctx.fetch("drainage_area"), ref["drainage"])
            assert np.array_equal(ctx.fetch("response_time"), ref["response"])


@pytest.mark.parametrize("name", ["uniform", "advanced", "uplift", "max_slope"])
def test_head_ordering_deep_nesting_path(oracle, emu_lib, name):
    """The head ordering of sweep 3 sorts on a fixed key base; segments nesting deeper than the base take a second
    pass with the exact base.  key_base=1 forces that path on small inputs."""
    m, p, outlets, initial, max_iteration = scenario(name)
    with _ctx(emu_lib, sweep=3, k5_split=0, key_base=1) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("opts", [dict(k5_cut=0), dict(k5_cut=1), dict(k5_cut=2), dict(k5_cut=4), dict(k5_cut=8),
                                  dict(k5_top_cap=0), dict(k5_top_cap=40), dict(k5_split=0)],
                         ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
@pytest.mark.parametrize("name", ["uniform", "advanced", "uplift", "max_slope", "mixed_slope", "plateau", "disconnected",
                                  "interior_outlets", "single_outlet", "lattice"])
def test_k5_split_by_nesting_height(oracle, emu_lib, name, opts):
    """K5 split at any cut height (everything through the run queue ... everything by per-height launches; the cut
    chosen from the level histogram with a tiny / zero queue budget) gives the oracle's bits; trees without an outlet
    stay unvisited."""
    m, p, outlets, initial, max_iteration = scenario(name)
    with _ctx(emu_lib, sweep=3, **opts) as ctx:
        helpers.load_ctx(ctx, m, p, outlets, initial)
        assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("name", ["uniform", "advanced", "uplift", "plateau", "disconnected", "interior_outlets"])
def test_first_iteration_levels_vs_flow(oracle, emu_lib, name):
    """first_flow=0: iteration 1 one launch per tree level; default: dataflow sweeps on a layout by subtree sizes."""
    m, p, outlets, initial, max_iteration = scenario(name)
    for first_flow in (0, 1):
        with _ctx(emu_lib, sweep=3, first_flow=first_flow) as ctx:
            helpers.load_ctx(ctx, m, p, outlets, initial)
            assert helpers.check_generate(ctx, oracle, m, p, outlets, initial, max_iteration)


@pytest.mark.parametrize("name,n", [("uniform", 2500), ("advanced", 4000), ("max_slope", 2000)])
def test_converged_properties_helper(oracle, emu_lib, name, n):
    """The size-independent property checks used by the full-size GPU tests (C2 at 1M, C3 at 4M sites), run here at a
    size where the oracle also converges, so the helper itself is known to hold on a correct result."""
    m, p, outlets, initial, _ = scenario(name, n)
    with _ctx(emu_lib) as ctx:
        e, it = helpers.check_converged_properties(ctx, oracle, m, p, outlets, initial)
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial)
    assert it == ref_it and np.array_equal(e, ref)


def test_ensemble_members_on_one_shared_graph(oracle, emu_lib):
    """C5-style ensemble on one model: members differ in erodibility, uplift, max_slope and (one of them) outlets; one
    context, one graph upload (fastlem_b200.ensemble.run_members_shared_graph).  The flood order survives a parameter
    change with the same outlets and is recomputed when the outlets change."""
    from fastlem_b200 import _native, ensemble
    from tools import workloads as W
    m, p, outlets, initial, _ = scenario("uniform", 2500)
    n = m["n"]
    rng = np.random.default_rng(4)
    other_outlets = np.unique(np.concatenate([outlets[::2], rng.integers(0, n, 5).astype(np.uint32)])).astype(np.uint32)
    members = [
        dict(initial=initial, erodibility=p["erodibility"], uplift=p["uplift"], outlets=outlets),
        dict(initial=initial, erodibility=0.5 + rng.random(n), uplift=p["uplift"], outlets=outlets),
        dict(initial=initial, erodibility=p["erodibility"], uplift=1.1 + 0.9 * W.value_noise(m["sites"], 0.05, seed=3, octaves=2),
             outlets=outlets),
        dict(initial=initial, erodibility=0.5 + rng.random(n), uplift=p["uplift"], outlets=other_outlets),
        dict(initial=initial, erodibility=p["erodibility"], uplift=p["uplift"], outlets=outlets,
             tan_max_slope=helpers.tan_of(np.full(n, 0.3))),
    ]
    res = ensemble.run_members_shared_graph(m, members, lambda: _native.Context(0, emu_lib), max_iteration=40)
    for k, (mem, (e, it)) in enumerate(zip(members, res)):
        ms = None if "tan_max_slope" not in mem else np.full(n, 0.3)
        ref, ref_it = oracle.generate(m, mem["erodibility"], mem["uplift"], ms, mem["outlets"], mem["initial"], 40)
        assert it == ref_it, k
        assert np.array_equal(e, ref), f"member {k}"


def test_graph_arrays_are_not_borrowed(oracle, emu_lib):
    """fastlem_set_graph copies its arrays (include/fastlem_b200.h): the caller may overwrite them right after the call,
    even when the flood order needs the exact host replay later on (tied edge lengths: the graph is read back from the
    device for it)."""
    m, p, outlets, initial, max_iteration = scenario("lattice_regular")
    with _ctx(emu_lib) as ctx:
        rp, col, dist, areas = (m[k].copy() for k in ("row_ptr", "col", "dist", "areas"))
        ctx.set_graph(rp, col, dist, areas)
        for a in ctx._keep["graph"]:
            a[...] = 0
        rp[...] = 0; col[...] = 0; dist[...] = 0; areas[...] = 0
        ctx.set_parameters(initial, p["erodibility"], p["uplift"], helpers.tan_of(p["max_slope"]), outlets)
        assert np.array_equal(ctx.fetch("flood_rank"), oracle.flood_order(m, outlets))
        assert ctx.stats()["flood_on_device"] == 0
        e, it = ctx.generate(max_iteration)
        ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, max_iteration)
        assert it == ref_it and np.array_equal(e, ref)


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("name", ["uniform", "uplift", "advanced"])
def test_speculative_receivers_leave_the_state_alone(oracle, emu_lib, name, overlap):
    """iterate_flow launches the NEXT iteration's K1 before the host has seen this iteration's flags (option `overlap`),
    into alternate buffers: after a run that stopped at its iteration cap or converged, the receivers, areas and
    response times that can be fetched are still those of the LAST iteration that counts."""
    m, p, outlets, initial, _ = scenario(name)
    for k in (2, 3, 6):
        before, _ = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, k - 1)
        ref = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, before)
        with _ctx(emu_lib, overlap=overlap) as ctx:
            helpers.load_ctx(ctx, m, p, outlets, initial)
            e, it = ctx.generate(k)
            assert it == k and np.array_equal(e, ref["elevations"])
            assert np.array_equal(ctx.fetch("receivers"), ref["next"])
            assert np.array_equal(ctx.fetch("drainage_area"), ref["drainage"])
            assert np.array_equal(ctx.fetch("response_time"), ref["response"])
    # ... and after convergence (the speculative K1 of the iteration that never runs has been launched)
    full, its = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, 400)
    if its < 400:
        last = oracle.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, full)
        with _ctx(emu_lib, overlap=overlap) as ctx:
            helpers.load_ctx(ctx, m, p, outlets, initial)
            e, it = ctx.generate(400)
            assert it == its and np.array_equal(e, full)
            assert np.array_equal(ctx.fetch("receivers"), last["next"])
