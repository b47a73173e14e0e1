import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build_libs(ref=False)
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    import pyoracle
    if os.path.isdir("/root/reference"):
        pyoracle.build_libs(ref=True)
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libhj3d_ref.so not built (no /root/reference here)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def pkg():
    import hj3d_loader
    return hj3d_loader.load()


@pytest.fixture(scope="session", params=["direct", "partitioned"])
def ctx(pkg, request):
    """Every GPU test runs twice: on the in-place path (small tables) and with bucket-range partitioning
    of build and probe inputs forced on (the path large tables take)."""
    import torch
    assert torch.cuda.is_available()
    # same stream as torch, so tensor fills / copies and engine kernels are ordered
    c = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    if request.param == "partitioned":
        c.set_option(pkg.OPT_PARTITION_BYTES, 1)
        c.set_option(pkg.OPT_PARTITION_WINDOW, 2048)
        c.set_option(pkg.OPT_PARTITION_MIN_PROBE, 0)
    c.mode = request.param
    return c
