import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu_lib():
    """Host emulation build of the solver (tests only; see fastlem_b200/csrc/fl_rt.h)."""
    from fastlem_b200 import build
    return build.build_emu()


@pytest.fixture(scope="session")
def product_lib():
    """The nvcc-built product library.  On the GPU box it is prebuilt and travels with the snapshot."""
    from fastlem_b200 import _native, build
    if not os.path.exists(_native.LIB_PATH):
        build.build()
    return _native.LIB_PATH


@pytest.fixture(scope="session")
def gpu_ctx_factory(product_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product has no CPU fallback")
    from fastlem_b200 import _native

    def make(**options):
        ctx = _native.Context(0, product_lib)
        for k, v in options.items():
            ctx.set_option(k, v)
        return ctx
    return make
