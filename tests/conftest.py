import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_cases():
    import glob
    import json
    out = []
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "jf_*.json"))):
        with open(p) as f:
            out.append(json.load(f))
    assert out, "no golden vectors"
    return out
