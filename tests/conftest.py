import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    """GPU context of the product library; gpu-marked tests fail (not skip) when it cannot be created."""
    from vpin_b200 import api

    c = api.Context(0)
    yield c
    c.close()
