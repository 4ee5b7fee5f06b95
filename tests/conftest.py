import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (incrementally) the product library and the oracle once per session."""
    import __graft_entry__ as ge

    ge.build()
    return True


@pytest.fixture(scope="session")
def cornell_small(tmp_path_factory, built):
    from btrc_b200 import scenes

    d = tmp_path_factory.mktemp("cornell_small")
    return scenes.make_cornell(str(d), width=96, height=96, spp=8, state_count=20000)


@pytest.fixture(scope="session")
def cornell_c1(tmp_path_factory, built):
    """BASELINE config 1 at full size: 512x512, 16 spp."""
    from btrc_b200 import scenes

    d = tmp_path_factory.mktemp("cornell_c1")
    return scenes.make_cornell(str(d), width=512, height=512, spp=16, state_count=1000000)
