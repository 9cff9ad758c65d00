import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.ljoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(density):
        return np.load(os.path.join(GOLDEN, "ref_%.1f.npz" % density))
    return load


def golden_rows(density):
    name = {0.5: "density0.5.dat", 1.0: "density1.dat"}[density]
    with open(os.path.join(GOLDEN, name)) as f:
        return [tuple(float(x) for x in line.split()) for line in f if line.strip()]
