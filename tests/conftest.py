import os
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
    """libbrq.so and the oracle binary; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    import helpers
    assert os.path.exists(helpers.ORACLE_CLI)
    return True


@pytest.fixture(scope="session")
def datasets(built, tmp_path_factory):
    """Synthetic BAM/FASTA pairs plus the oracle's outputs for each, generated once per session."""
    import helpers
    root = tmp_path_factory.mktemp("brq_data")
    out = {}
    for name in helpers.DATASETS:
        out[name] = helpers.make_dataset(name, str(root / name))
    return out
