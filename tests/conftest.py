"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` covers the oracle (against the reference's golden vectors), the host logic and the
C-ABI's symbol table; `-m gpu` holds the parity tests proper and runs on a B200 through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box only)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import feo_oracle
    feo_oracle.lib()
    return feo_oracle


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")) as f:
        return json.load(f)["cases"]
