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


@pytest.fixture(scope="session", params=["direct", "partitioned", "smem", "default"])
def ctx(pkg, request):
    """Every GPU test runs on four engine configurations: in place with global-memory lookups (small inputs),
    bucket-range partitioned with L2-window lookups, partitioned into fine partitions probed in shared memory
    (what large inputs take; the latter two forced on at test sizes), and the default options with the
    compressed-slice probe of large unique-key probes forced on at test sizes (tests/test_gpu_parity_large.py runs the
    untouched defaults at full size)."""
    import torch
    assert torch.cuda.is_available()
    # same stream as torch, so tensor fills / copies and engine kernels are ordered
    c = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    if request.param == "direct":
        c.set_option(pkg.OPT_SMEM_PROBE, 0)
        c.set_option(pkg.capi.OPT_SMEM_BUILD, 0)
    if request.param == "partitioned":
        c.set_option(pkg.OPT_SMEM_PROBE, 0)
        c.set_option(pkg.capi.OPT_SMEM_BUILD, 0)
        c.set_option(pkg.OPT_PARTITION_BYTES, 1)
        c.set_option(pkg.OPT_PARTITION_WINDOW, 2048)
        c.set_option(pkg.OPT_PARTITION_MIN_PROBE, 0)
        c.set_option(pkg.capi.OPT_PART_THREADS, 256)       # also cover the other partition kernel variants
        c.set_option(pkg.capi.OPT_PART_RANK_MATCH, 1)
    if request.param == "smem":
        c.set_option(pkg.OPT_PARTITION_BYTES, 1)
        c.set_option(pkg.OPT_PARTITION_WINDOW, 65536)
        c.set_option(pkg.OPT_SMEM_MIN_PROBE, 0)
        c.set_option(pkg.OPT_SMEM_SLICE_BYTES, 4096)
        c.set_option(pkg.capi.OPT_SMEM_BUILD_BYTES, 4096)
        c.set_option(pkg.OPT_SMEM_CHUNK, 4096)
        c.set_option(pkg.capi.OPT_PART_SAMPLE, 2)            # regions planned from a sampled histogram (what skewed inputs take)
    if request.param == "default":
        # default options, except that the compressed-slice probe (what 2^22+ row probe sides take) is forced on with
        # small slices, so that several fine partitions, two partition levels and slices that do not fit all occur
        c.set_option(pkg.capi.OPT_PACKED_PROBE, 1)
        c.set_option(pkg.capi.OPT_PACKED_MIN_PROBE, 0)
        c.set_option(pkg.capi.OPT_PACKED_SLICE_BYTES, 4096)
    c.mode = request.param
    return c
