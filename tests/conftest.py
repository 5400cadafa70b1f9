import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import surf_oracle

    surf_oracle.lib()
    return surf_oracle


@pytest.fixture(scope="session")
def frame640():
    from surffeatures_b200 import speckle

    return speckle.speckle_frame(480, 640, 0)
