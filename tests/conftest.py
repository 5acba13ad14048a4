import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a CUDA device: on a machine without one they are skipped (not errored), so a plain
    `pytest tests` is green on CPU-only CI.  (The product path itself never falls back: it fails loudly.)"""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle(pkg):
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def shim(pkg):
    import shim_lib
    shim_lib.lib()
    return shim_lib


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context([0])
    yield c
    c.close()
