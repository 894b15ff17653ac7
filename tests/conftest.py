import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libb2sense.so if the tree arrived without it (nvcc cross-compiles without a GPU)."""
    try:
        from deep_cine_cardiac_mri_b200 import _lib
        if not _lib.LIB_PATH.exists():
            _lib.build()
    except Exception as e:                       # the tests that need it will fail loudly on their own
        print("could not build libb2sense.so:", e)
    yield
