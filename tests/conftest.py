import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Make sure the product library exists (a fresh checkout has no built artefacts; nvcc cross-compiles
    without a GPU).  No-op when mpt_b200/_lib/libmptg.so is up to date."""
    try:
        from mpt_b200 import build

        build.build()
    except Exception as e:  # the tests that need the library will fail loudly on their own
        print(f"warning: could not build libmptg.so: {e}", file=sys.stderr)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand from oracle/."""
    from tests import oracle_binding

    return oracle_binding.load()


@pytest.fixture(scope="session")
def ctx():
    import mpt_b200

    c = mpt_b200.Context(0)
    yield c
    c.close()
