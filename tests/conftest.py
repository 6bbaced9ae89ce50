import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need the CUDA engine: on a box without a device they are skipped (a plain `pytest tests` stays green), and a
    stale libggnn_b200.so that could not be rebuilt is reported instead of being used silently."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if not has_cuda:
        skip = pytest.mark.skip(reason="needs a CUDA device (B200): run with -m gpu on the GPU box")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
