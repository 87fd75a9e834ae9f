import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# Small dense factorisations dominate the oracle; with one BLAS thread per core and anything else running on the machine
# the threads spin on each other (measured: 200 s instead of 4 s for one oracle level).  A few threads are enough.
try:
    from threadpoolctl import threadpool_limits
    _blas_limit = threadpool_limits(limits=int(os.environ.get("ALFIB_TEST_BLAS_THREADS", "4")))
except Exception:       # noqa: BLE001 - threadpoolctl is optional
    _blas_limit = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def problems():
    """Cache of small synthetic problems shared by the tests."""
    from alfi_b200.synth.problem import build_problem
    cache = {}

    def get(name, **kw):
        key = (name, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = build_problem(name, **kw)
        return cache[key]
    return get
