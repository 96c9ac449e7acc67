import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pgo_oracle

    pgo_oracle.build()
    pgo_oracle.lib()
    return pgo_oracle


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through the C-ABI.  Fails loudly (no skip, no CPU fallback) if unusable."""
    from pose_graph_initialization_b200 import Engine

    e = Engine(device=0)
    yield e
    e.close()
