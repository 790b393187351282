import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcdyn.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests get a hard per-test time limit (pytest-timeout, when installed): a kernel that spins — a peer that never
    posts its flag, a ring that is never drained — must cost one test, not the GPU box's whole time slot."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(120))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces in-tree if they are missing or stale (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build_cuda()
    ge.build_oracle()
    yield


@pytest.fixture(scope="session")
def ctx():
    import gcdyn
    return gcdyn.Context.default()
