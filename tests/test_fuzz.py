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
