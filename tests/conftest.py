import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long CPU oracle runs (enabled by LBM_FULL_GOLDEN=1)")


@pytest.fixture(scope="session")
def lbm():
    """the ctypes view of the product library (hyphenated package name -> importlib)"""
    return importlib.import_module("hpc-lattice-boltzmann_b200")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """build the checkers/products once per session if something is missing (CPU only: nvcc and
    gcc cross-compile without a GPU).  On the GPU box everything arrives prebuilt."""
    need = [os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "oracle", "canon"),
            os.path.join(ROOT, "hpc-lattice-boltzmann_b200", "liblbm_b200.so"),
            os.path.join(ROOT, "hpc-lattice-boltzmann_b200", "d2q9-bgk.exe")]
    if not all(os.path.isfile(p) for p in need):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
