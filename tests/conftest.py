import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import pkg_loader
    return pkg_loader.load()


@pytest.fixture(scope="session")
def O():
    """The CPU oracle + reference window (test infrastructure)."""
    import oracle_py
    oracle_py.build()
    return oracle_py


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def unhex(s):
    return np.frombuffer(bytes.fromhex(s), dtype=np.float64).copy()


def unhex1(s):
    return float(unhex(s)[0])


@pytest.fixture(scope="session")
def golden_ecc():
    return load_golden("ecc_kats.json")


@pytest.fixture(scope="session")
def golden_cg():
    return load_golden("cg_poisson2d_16.json")


@pytest.fixture(scope="session")
def golden_cli():
    return load_golden("cli_poisson2d_32.json")
