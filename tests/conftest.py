import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "kats.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def small_case():
    """A seeded prior forest + points small enough for the numpy oracle (seconds)."""
    from bartint_b200 import synth
    rng = np.random.default_rng(np.random.PCG64(11))
    x_train = rng.random((40, 4))
    y_train = synth.genz_gaussian(x_train)
    ff = synth.prior_forest(5, 12, 9, x_train, y_train)
    X = rng.random((3001, 4))
    return ff, x_train, y_train, X
