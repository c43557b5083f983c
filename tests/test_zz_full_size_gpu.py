"""GPU tier, run last (file name): BASELINE configurations at their full sizes, checked through size-independent
properties because the oracle cannot run them to convergence in test time (tests/helpers.py,
check_converged_properties; the helper itself is validated on the CPU tier in test_emu_parity.py).
C2 at 1M sites lives in test_gpu_parity.py::test_c2_one_million_sites."""
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
