import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle and the synthetic-data tool once per session (CPU only, seconds)."""
    from matchtigs_b200 import _build
    _build.build_oracle()
    _build.build_synth()
    yield
