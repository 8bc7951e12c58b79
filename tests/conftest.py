"""Test configuration. `-m "not gpu"` runs on a CPU-only box (oracle vs goldens, host logic,
C-ABI symbol checks); `-m gpu` needs a B200 and drives the CUDA path through the C ABIs."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import _pkg  # noqa: E402



class _Kit:
    """What the tests call `pkg`: the product package plus, under `.oracle`, the loader of the CPU
    oracle (oracle/facade.py). The product package itself knows nothing about the oracle."""

    def __init__(self, package, oracle):
        self._package, self.oracle = package, oracle

    def __getattr__(self, name):
        return getattr(self._package, name)


trg = _Kit(_pkg.load(), _pkg.load_oracle().oracle)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pkg():
    return trg


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle and the product libraries exist (builds them if missing)."""
    import __graft_entry__ as g
    g.build(quiet=True, only_missing=True)
    return True


def has_gpu() -> bool:
    try:
        from importlib import import_module
        k = import_module("trg_planner_b200.kernels")
        return k.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def small_mountain(pkg):
    return pkg.terrain.mountain(300, h=0.1, seed=2)


@pytest.fixture(scope="session")
def small_indoor(pkg):
    return pkg.terrain.indoor(150, h=0.2, seed=1)
