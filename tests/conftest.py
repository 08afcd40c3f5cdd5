import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "query-engines_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libkqoracle.so) — the checker, never the thing shipped."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu():
    """The product: libkqgpu.so through the Python mirror of the reference's operator API."""
    import kqgpu
    if kqgpu.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box (there is no CPU fallback)")
    return kqgpu


@pytest.fixture(scope="session")
def gctx(gpu):
    return gpu.Context(0)
