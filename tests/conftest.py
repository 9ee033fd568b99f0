import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    try:
        from xvc_b200 import lib
        return lib.device_count()
    except Exception:  # noqa: BLE001
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a CUDA device skips the GPU tests instead of
    failing them (the oracle / reference parity tests stay visible).  When GPU tests are asked for
    explicitly (`-m gpu`, the driver's run on the B200 box) or XVCB_REQUIRE_GPU=1 is set, nothing is
    skipped: a missing device or extension then fails loudly."""
    markexpr = (config.getoption("markexpr") or "").replace(" ", "")
    explicit = "gpu" in markexpr and "notgpu" not in markexpr
    if explicit or os.environ.get("XVCB_REQUIRE_GPU") or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
