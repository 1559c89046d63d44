"""pytest configuration: registers the `gpu` marker and puts the product package and the oracle on sys.path.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol check (no compute calls).
`-m gpu`       : parity tests proper -- CUDA path through the C-ABI vs the oracle / golden vectors.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
