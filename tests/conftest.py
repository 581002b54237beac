import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    g = {}
    for name in ("reference_vectors", "exact_binom", "tree_vectors", "basket_vectors"):
        with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
            g[name] = json.load(f)
    return g


@pytest.fixture(scope="session")
def gpu():
    """libpcf.so initialised on cuda:0. Fails (does not skip) if the extension is missing."""
    import parcompfin_b200 as pcf
    pcf.load_library()
    pcf.init(1)
    yield pcf
    pcf.shutdown()
