import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (builder container only)")


def pytest_collection_modifyitems(config, items):
    from oracle.ref_loader import reference_available
    if reference_available():
        return
    skip = pytest.mark.skip(reason="/root/reference not present (GPU box): golden fixtures cover this")
    for item in items:
        if "reference" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
