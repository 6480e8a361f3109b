"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"`: oracle vs its pins, host logic, library load / symbol export (runs anywhere).
`-m gpu`      : parity tests proper — the CUDA path through the C ABI vs the oracle (B200).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libgpsurv.so")


@pytest.fixture(scope="session")
def lib():
    from gpsurvival_b200 import build, _lib
    build.build_lib()
    return _lib.load()
