import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import bindings
    return bindings.Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import bindings
    if not bindings.have_ref():
        pytest.skip("oracle/_ref/libxvcref.so not built (needs /root/reference)")
    return bindings.Ref()
