import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rl-ode-physics_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(autouse=True)
def _guard_bands(request):
    """With ODE_B200_DEBUG_GUARD=1 (tests/test_guards_gpu.py re-runs GPU tests that way) every device allocation of
    the library carries guard bands; a test that made a kernel write outside an allocation fails here."""
    yield
    if os.environ.get("ODE_B200_DEBUG_GUARD") == "1" and request.node.get_closest_marker("gpu"):
        import odeb200
        bad = odeb200.lib().dCheckGuardsB200(1)
        assert bad == 0, "%d device allocation(s) with a damaged guard band (named on stderr)" % bad
