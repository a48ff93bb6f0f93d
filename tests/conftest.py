"""pytest configuration: `gpu` marker, repo root on sys.path, golden fixture loader."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the reference install (git-ignored; only present when baseline/install_ref.sh has been run)
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(_REF, "qibo")) and _REF not in sys.path:
    sys.path.append(_REF)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """tests/golden/reference_golden.npz -- outputs of the unmodified reference NumpyBackend."""

    def __init__(self):
        self.z = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))

    def cases(self, key):
        return json.loads(str(self.z[key]))

    def __getitem__(self, key):
        return self.z[key]

    def __contains__(self, key):
        return key in self.z.files


@pytest.fixture(scope="session")
def golden():
    return Golden()


def have_qibo():
    try:
        import qibo  # noqa: F401

        return True
    except Exception:
        return False


def tol(dtype):
    """north_star tolerances: 1e-12 max-abs for complex128, 1e-5 for complex64."""
    return 1e-12 if str(dtype) in ("complex128", "float64") else 1e-5
