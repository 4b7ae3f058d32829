import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import gp_oracle
    return gp_oracle


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_ds(oracle, gold, n):
    return oracle.make_inference_datasets(gold["X"][:n], gold["Y"][:n], gold[f"hyp_{n}"])


@pytest.fixture(scope="session")
def c1():
    return load_golden("c1_benoit")


@pytest.fixture(scope="session")
def c3():
    return load_golden("c3_wor")


@pytest.fixture(scope="session")
def engine():
    """One GridEngine (one sbo_ctx on cuda:0) for the GPU session; fails loudly when no GPU / no .so."""
    import sbo_b200
    eng = sbo_b200.GridEngine(0)
    yield eng
    eng.close()
