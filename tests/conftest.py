import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = {
    "data_chr1": os.path.join(GOLDEN, "data_chr1", "data_chr1"),   # flashpcaR/inst/extdata
    "hapmap3": os.path.join(GOLDEN, "hapmap3", "data"),            # HapMap3/data (BASELINE cfg 0)
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fixtures():
    return FIXTURES


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load the native library; never falls back."""
    from flashpca_b200 import _lib, build
    build.build_lib()
    return _lib.load()


def load_fixture(name):
    from oracle import oracle as O
    stem = FIXTURES[name]
    n = O.count_lines(stem + ".fam")
    payload, npb, p = O.read_bed_payload(stem + ".bed", n)
    return stem, payload, n, p
