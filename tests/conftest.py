import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): C restatement of the BMAS contract."""
    import oracle as O
    return O.load()


@pytest.fixture(scope="session")
def refnu():
    from oracle import reference_nu
    return reference_nu


@pytest.fixture(scope="session")
def nu():
    """The product package on cuda:0.  Fails loudly (no fallback) when the device is missing."""
    import numericals_b200 as nu
    nu.require_device(0)
    return nu


@pytest.fixture(scope="session")
def libnuk():
    from numericals_b200 import _lib
    return _lib.lib()
