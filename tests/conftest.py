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


@pytest.fixture(scope="session")
def engine_variants():
    """An engine on the TEST build of the library (tests/native/libswb200_variants.so, -DSWB_ALL_VARIANTS): every
    short-read kernel variant, cross-checked against the oracle.  The product library carries the default only."""
    import mini_parallel_b200 as mp
    from mini_parallel_b200 import _lib
    path = os.path.join(ROOT, "tests", "native", "libswb200_variants.so")
    if not os.path.exists(path):
        pytest.fail("tests/native/libswb200_variants.so is missing: make variants (or __graft_entry__.build())")
    eng = mp.Engine(0, lib=_lib.bind(path))
    yield eng
    eng.close()


@pytest.fixture(autouse=True)
def _arena_guards(request):
    """SWB_GUARD=1 (tools/sanitize.sh): after every GPU test, no guard zone round any device arena may have been written."""
    yield
    if os.environ.get("SWB_GUARD", "0") in ("", "0") or "engine" not in request.fixturenames:
        return
    eng = request.getfixturevalue("engine")
    bad, n, report = eng.guard_check()
    assert bad == 0 and n > 0, f"{bad} damaged guard zones round {n} arenas:\n{report}"
    if "engine_variants" in request.fixturenames:
        bad, n, report = request.getfixturevalue("engine_variants").guard_check()
        assert bad == 0, report
