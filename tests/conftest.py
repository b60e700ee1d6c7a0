import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the tests bind the in-tree CUDA library; build it if a fresh checkout has none yet (nvcc cross-compiles without a
    # GPU).  A stale-looking timestamp alone does not trigger a rebuild: __graft_entry__.build() is the build step.
    try:
        from hhmarl_2d_b200 import _native as nat
        if not os.path.exists(nat.LIB_PATH):
            nat.build()
    except Exception as e:  # noqa: BLE001 -- the tests that need the library will say so themselves
        print(f"[conftest] could not build {e!r}")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
