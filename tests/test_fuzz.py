"""CPU tier: differential fuzz of the solver (host emulation build, every sweep implementation) against the oracle on
random small graphs that no scenario covers: non-planar, random degrees up to 32, tied edge lengths, several components,
random outlets, mixed max_slope, non-uniform uplift, random iteration caps.  Everything bit-exact.
(tools/fuzz_solver.py runs the same generator over thousands of seeds and larger graphs.)"""
import numpy as np
import pytest

import helpers
from tools.fuzz_solver import random_case


@pytest.mark.parametrize("block", range(6))
def test_random_graphs_bit_exact(oracle, emu_lib, block):
    from fastlem_b200 import _native
    done = 0
    for seed in range(block * 12, block * 12 + 12):
        case = random_case(seed, oracle)
        if case is None:
            continue
        m, p, outlets, initial, max_iteration = case
        ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, max_iteration)
        for sweep in (0, 1, 2, 3):
            with _native.Context(0, emu_lib) as ctx:
                ctx.set_option("sweep", sweep)
                helpers.load_ctx(ctx, m, p, outlets, initial)
                e, it = ctx.generate(max_iteration)
            assert it == ref_it, (seed, sweep)
            assert np.array_equal(e, ref, equal_nan=True), (seed, sweep)
        done += 1
    assert done >= 6


@pytest.mark.parametrize("seed,opts", [
    (20758, dict(incremental=1, incr_div=16, rebuild_every=5, park_after=4, key_base=2, fuse_levels=1, first_flow=0,
                 flood_device=0, rebuild_growth=4, rebuild_height=150)),
    (20332, dict(incremental=1, incr_div=1, rebuild_every=5, park_after=8, key_base=3, fuse_levels=0, first_flow=0,
                 flood_device=0, rebuild_growth=50, rebuild_height=400)),
])
def test_height_bound_above_key_base_regression(oracle, emu_lib, seed, opts):
    """Found by tools/fuzz_solver.py --options: in an incremental iteration the bound on the nesting height carried over
    from the previous iteration can lie above the key base although no key of this iteration overflowed it; the head
    ordering then has to be redone with the exact base (it used to fail with 'height bookkeeping broke')."""
    from fastlem_b200 import _native
    m, p, outlets, initial, mi = random_case(seed, oracle, 2, 150, (1, 40))
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, mi)
    with _native.Context(0, emu_lib) as ctx:
        for k, v in opts.items():
            ctx.set_option(k, v)
        helpers.load_ctx(ctx, m, p, outlets, initial)
        e, it = ctx.generate(mi)
    assert it == ref_it and np.array_equal(e, ref, equal_nan=True)


def test_context_reuse_sequences(oracle, emu_lib):
    """One context through random sequences of set_graph / set_parameters / set_option('sweep') / generate: graphs of
    different sizes, outlets that change, max_slope present in one run and absent in the next (this used to leave the
    level-synchronous path clamping with the previous run's slopes), plateaus; every run bit-exact against the oracle."""
    from fastlem_b200 import _native
    from tools.fuzz_solver import random_graph
    for seed in range(25):
        rng = np.random.default_rng(seed)
        graphs = [random_graph(rng, 5, 200) for _ in range(2)]
        with _native.Context(0, emu_lib) as ctx:
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


def check_special_parameter_values(lib, oracle, seeds):
    """Zero, negative, infinite and denormal-range erodibility / uplift on a tenth of the sites: infinities and NaNs
    appear in response times and elevations and must propagate exactly as in the oracle (f64::max semantics of
    generator.rs:179, IEEE division), for the level-synchronous and the dataflow sweeps."""
    from fastlem_b200 import _native
    for seed in seeds:
        m, p, outlets, initial, mi = random_case(seed, oracle, 2, 120, (1, 30))
        rng = np.random.default_rng(seed + 9)
        n = m["n"]
        k, u = p["erodibility"].copy(), p["uplift"].copy()
        idx = rng.choice(n, max(1, n // 10), replace=False)
        kind = seed % 6
        if kind == 0:
            k[idx] = 0.0
        elif kind == 1:
            k[idx] = 1e300
        elif kind == 2:
            u[idx] = 0.0
        elif kind == 3:
            u[idx] = -1.0
        elif kind == 4:
            k[idx] = np.inf
        else:
            u[idx] = 1e-300
            k[idx] = 1e-300
        ref, ref_it = oracle.generate(m, k, u, p["max_slope"], outlets, initial, mi)
        for sweep in (0, 3):
            with _native.Context(0, lib) as ctx:
                ctx.set_option("sweep", sweep)
                ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
                ctx.set_parameters(initial, k, u, helpers.tan_of(p["max_slope"]), outlets)
                e, it = ctx.generate(mi)
            assert it == ref_it and np.array_equal(e, ref, equal_nan=True), (seed, kind, sweep)


def test_special_parameter_values(oracle, emu_lib):
    check_special_parameter_values(emu_lib, oracle, range(60))
