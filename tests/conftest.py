import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """ctypes handle on oracle/libpapr_oracle.so (the CPU restatement; test infrastructure)."""
    import oracle_binding
    return oracle_binding.load()


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the product library + CLI; returns the python package."""
    import __graft_entry__
    __graft_entry__.build()
    import dtv_utils_b200
    return dtv_utils_b200
