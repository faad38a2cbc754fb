import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly on a box without a GPU unless deselected with -m 'not gpu'."""
    return


@pytest.fixture(scope="session")
def technical_golden():
    with open(os.path.join(GOLDEN_DIR, "technical_golden.json")) as f:
        return json.load(f)


def approx_rel(a, b, rel=1e-3, abs_=1e-9):
    import math
    if isinstance(a, float) and isinstance(b, float):
        if math.isinf(a) or math.isinf(b) or math.isnan(a) or math.isnan(b):
            return (math.isnan(a) and math.isnan(b)) or a == b
    return abs(a - b) <= max(abs_, rel * max(abs(a), abs(b)))
