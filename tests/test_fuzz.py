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
