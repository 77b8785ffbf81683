import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def synth():
    import vloam_b200  # noqa: F401  (registers the package)
    from vloam_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def scans_small(synth):
    """4 consecutive scans of seed 7 at 512 columns (64 x 512 = 32768 points) — fast enough for the CPU suite."""
    return synth.make_scans(7, 4, n_cols=512)


@pytest.fixture(scope="session")
def scans_full(synth):
    """3 consecutive full-size scans (64 x 2048) of the benchmark seed."""
    return synth.make_scans(1234, 3)
