import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def lib():
    """The product library; built on demand where nvcc exists, never substituted."""
    from leann_b200 import build, capi

    if build.needs_build():
        build.build()
    return capi.load()


@pytest.fixture(scope="session")
def cuda_ok():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return True
