import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _pin_cuda_device_count():
    """The reference's select_device("cpu") sets CUDA_VISIBLE_DEVICES=-1 in this process (utils/torch_utils.py:176); the first
    torch.cuda.device_count() after that would cache 0 for good and a later device=0 call of the SAME reference function
    would be refused.  Cache the true count while the environment is still intact (tests mix CPU reference runs with ours)."""
    import torch

    if torch.cuda.is_available():
        torch.cuda.init()
        torch.cuda.device_count()
    yield
