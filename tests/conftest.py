import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cuda_lib():
    """The CUDA library, loaded; GPU tests must run the native path or fail loudly."""
    import torch
    from evfly_b200 import _lib
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    lib = _lib.load()
    cap = torch.cuda.get_device_capability()
    assert cap[0] == 10, f"libevfly_b200 is built for sm_100a only, found sm_{cap[0]}{cap[1]}"
    return lib
