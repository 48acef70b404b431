"""pytest configuration: marker registration and shared fixtures."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class Golden:
    """Accessor for tests/golden/post_golden.npz (outputs of the reference's own functions)."""

    def __init__(self, path):
        self.z = np.load(path)
        self.index = json.loads(bytes(self.z["cases"]).decode())

    def __getitem__(self, k):
        return self.z[k]

    def meta(self, k):
        return json.loads(bytes(self.z[k]).decode())


@pytest.fixture(scope="session")
def golden():
    return Golden(ROOT / "tests" / "golden" / "post_golden.npz")
