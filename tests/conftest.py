import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _native_libraries_built():
    """A fresh checkout has no .so files (they are git-ignored): build the CUDA library
    (nvcc cross-compiles without a GPU) and the CPU oracle once per session if missing."""

    from reinfocus_b200 import build as native_build

    if not os.path.exists(native_build.LIB_PATH):
        native_build.build()
    import oracle

    oracle.build()
