import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "first_gpu_run: GPU leg written after the round's GPU minutes were spent — its "
                            "host logic is CPU-checked, the hardware run is still pending; reported as xfail/xpass "
                            "(non-strict) so that an unverified leg cannot turn the verified suite red. "
                            "B200GF_STRICT_WIDEN=1 (tools/gpu_first_check.sh) makes these ordinary tests.")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if os.environ.get("B200GF_STRICT_WIDEN", "0") != "1":
        pending = pytest.mark.xfail(strict=False, reason="first hardware run pending (see marker first_gpu_run)")
        for item in items:
            if "first_gpu_run" in item.keywords:
                item.add_marker(pending)
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
