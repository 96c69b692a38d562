"""pytest configuration.

Markers:
  gpu -- needs a CUDA device (run on the B200 box: ``pytest -m gpu``); everything
         else runs on a CPU-only machine (``pytest -m "not gpu"``).

The CPU suite checks the oracle against golden vectors, the host logic and the
kernel arithmetic compiled for the host; the GPU suite is the parity suite
proper and goes through the C ABI (include/gidcuda.h).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_IMG = "/root/reference/test/img"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def oracle():
    from tests.harness import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def synth():
    from tests.harness import Synth
    return Synth()


@pytest.fixture(scope="session")
def emu():
    from tests.harness import Emu
    return Emu()


@pytest.fixture(scope="session")
def gidcuda_lib():
    """The product library; built on demand (in-tree), never replaced by a fallback."""
    from gid_b200 import build, capi
    if not os.path.exists(capi.library_path()):
        build.build_cuda()
    return capi.load_library()


@pytest.fixture(scope="session")
def gpu(gidcuda_lib):
    """A context on cuda:0.  Fails (does not skip) when there is no device:
    GPU tests must never pass on a fallback."""
    from gid_b200 import GidCuda
    ctx = GidCuda(0)
    yield ctx
    ctx.close()


def reference_image(name: str) -> str:
    path = os.path.join(REFERENCE_IMG, name)
    if not os.path.exists(path):
        pytest.skip("reference fixtures are not present on this machine")
    return path
