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
def engine():
    import mini_parallel_b200 as mp
    if mp.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible (there is no CPU fallback)")
    eng = mp.Engine(0)
    yield eng
    eng.close()
