import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
