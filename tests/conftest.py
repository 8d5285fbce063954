import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "recovery-rl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def native():
    """The ctypes binding, with librrl.so built if it is missing (build is CPU-only: nvcc cross-compiles)."""
    lib_path = os.path.join(PKG, "librrl.so")
    if not os.path.exists(lib_path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("rrl_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    from recovery_rl import native as n
    return n


@pytest.fixture(scope="session")
def cuda(native):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
