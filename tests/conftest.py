import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build the CUDA library
    in-tree (nvcc cross-compiles without a GPU, ~1 min) and the oracle before collecting, so that
    `pytest tests` works without a separate build step.  Nothing is built if it already exists."""
    from microfc_b200 import abi
    if not os.path.exists(abi.LIB_PATH) and not os.environ.get("MFC_B200_LIB"):
        try:
            from microfc_b200 import build
            build.build()
        except Exception as e:                       # the tests that need the library will fail loudly
            print(f"[conftest] could not build libmfc_b200.so: {e}", file=sys.stderr)
