"""pytest configuration: the `gpu` marker, import paths, and shared fixtures.

    python -m pytest tests -x -q -m "not gpu"     CPU-only: oracle vs known answers / golden fixtures, host logic, ABI surface
    python -m pytest tests -x -q -m gpu           on a B200: the parity tests proper, through the C ABI (libycge.so)
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(TESTS, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything is built in-tree once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle_binding import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def pkg():
    import yetanotherconsolegameengine_b200 as p
    return p


def require_cuda():
    """GPU tests FAIL (not skip) when the CUDA path is unusable: there is no CPU fallback to hide behind."""
    import torch
    assert torch.cuda.is_available(), "a test marked `gpu` ran without a CUDA device"
