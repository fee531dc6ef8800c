import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU parity oracle (oracle/liboracle.so), built on demand.  Checker only."""
    from oracle.loader import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own code compiled IEEE-strict (oracle/_ref), when it has been built."""
    from oracle.loader import load_reference, reference_available
    if not reference_available("strict"):
        pytest.skip("oracle/_ref/libpda_ref_strict.so not built (needs /root/reference: make -C oracle ref)")
    return load_reference("strict")


@pytest.fixture(scope="session")
def gpu_api():
    """The product: probabilisticsemslam_b200.api over libpda_b200.so.  Fails loudly without a device."""
    from probabilisticsemslam_b200 import api, _lib
    assert _lib.lib().pda_device_count() > 0, "no CUDA device visible: -m gpu tests must run on the GPU box"
    return api


@pytest.fixture(params=["warp", "cta", "fast"])
def murty_path(request, gpu_api):
    """Runs a test once per Murty kernel: one warp per problem with the exact heap ("warp"), one CTA per problem
    (latency, "cta": applies wherever numCol <= 16, the warp kernel takes the rest) and the pruning kernel with its
    exact fallback for tied problems ("fast": the default for large batches).  All must give the reference's bits."""
    prev = gpu_api.set_murty_path(request.param)
    yield request.param
    gpu_api.set_murty_path(prev)
