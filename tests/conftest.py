import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the upstream tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    class _G(dict):
        @property
        def files(self):
            return list(self.keys())

    def _load(name):
        if name not in cache:
            with np.load(os.path.join(GOLDEN, name)) as z:
                cache[name] = _G({k: z[k] for k in z.files})  # decompress once
        return cache[name]
    return _load


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
