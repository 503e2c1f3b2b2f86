import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def vsb():
    import vsb200_loader

    return vsb200_loader.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def gpu_vsb(vsb):
    """The product library on a real device. Fails loudly (never skips to a fallback) when unusable."""
    if not os.path.exists(vsb.LIB_PATH):
        pytest.fail("libvsb200.so is not built: run __graft_entry__.build()")
    if vsb.device_count() < 1:
        pytest.fail("no CUDA device visible to libvsb200 (GPU tests must run on the GPU box)")
    return vsb
