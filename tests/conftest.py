import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    if os.environ.get("XDE_DRY_RUN") == "1":  # host-layer dry run without a GPU: see tests/host_dry_run.py
        from tests import host_dry_run

        config._xde_dry_run = host_dry_run.install()


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (test infrastructure; built on demand with gcc)."""
    from oracle import xde_oracle

    xde_oracle.build()
    return xde_oracle
