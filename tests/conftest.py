"""pytest configuration: `gpu` marker (needs a B200 + built libeemflow_b200.so) and shared helpers."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected without a device: a silent skip would hide a
    missing CUDA path.  Without `-m gpu` they are simply deselected by the marker expression."""
    return


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(GOLDEN / f"{name}.npz")
    return load


def cases_of(npz, suffix="__events"):
    return sorted(k[: -len(suffix)] for k in npz.files if k.endswith(suffix))
