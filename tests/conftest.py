import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand where nvcc exists)."""
    from hippomm_b200 import _lib, build

    if not _lib.LIB_PATH.exists():
        build.build_library()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device(lib):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from hippomm_b200 import _cuda

    return _cuda.require_device()
