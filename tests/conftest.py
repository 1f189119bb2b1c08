"""pytest configuration: markers and shared fixtures."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    """Compile the CPU checkers once per session (C restatement; reference engine if its sources exist)."""
    from oracle import int_oracle
    int_oracle.build()
    return int_oracle
