import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    """Factory for lmono_b200 contexts; fails loudly (no CPU fallback) if CUDA is unusable."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from lmono_b200 import api

    made = []

    def make(**params):
        c = api.Context(device=0, **params)
        made.append(c)
        return c

    make.made = made
    yield make
    for c in made:
        c.close()


@pytest.fixture(autouse=True)
def _close_contexts_of_this_test(request):
    """A ctx holds ~3 GB of device memory (slab pool of the cube map).  Contexts a test creates itself are closed when
    the test ends; those of module-scoped fixtures (set up before this function-scoped fixture runs) live on."""
    if "gpu_ctx_factory" not in request.fixturenames:
        yield
        return
    fac = request.getfixturevalue("gpu_ctx_factory")
    n0 = len(fac.made)
    yield
    for c in fac.made[n0:]:
        c.close()
    del fac.made[n0:]
