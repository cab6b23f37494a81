import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle tools and the CUDA library exist (cross-compiles without a GPU)."""
    import helpers
    helpers.ensure_built()
    return True


@pytest.fixture(scope="session")
def workspace(built, tmp_path_factory):
    """Genomes, indexes (built with the reference binary) and simulated reads."""
    import helpers
    if not os.path.exists(helpers.REF_BIN):
        pytest.skip("oracle/_ref/abismal not built (needs /root/reference once; see oracle/Makefile)")
    d = str(tmp_path_factory.mktemp("ws"))
    return helpers.Workspace(d)
