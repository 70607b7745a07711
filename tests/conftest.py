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
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def crt_lib():
    from computational_ray_tracer_b200 import build as _build
    _build.build()
    from computational_ray_tracer_b200 import _capi
    return _capi.load()


@pytest.fixture(scope="session")
def gpu_ctx(crt_lib):
    from computational_ray_tracer_b200 import api
    ctx = api.Context(0)      # raises without a CUDA device: there is no CPU fallback
    yield ctx
    ctx.close()
