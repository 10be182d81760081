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
    """Loader for tests/golden fixtures (produced by the reference's own code, oracle/make_golden.py)."""
    import numpy as np
    cache = {}

    def load(kind, name):
        key = (kind, name)
        if key not in cache:
            with np.load(os.path.join(ROOT, "tests", "golden", "%s_%s.npz" % (kind, name))) as z:
                cache[key] = {k: z[k] for k in z.files}
        return cache[key]
    return load
