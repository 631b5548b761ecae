import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


def pytest_collection_modifyitems(config, items):
    """A green `-m gpu` run must mean the CUDA path ran on a GPU: the CPU emulation build of tools/emu (selected with
    B2R_LIBRARY) is test infrastructure and may stand in for the GPU only when tests/test_cpu_emu.py asks for it
    (B2R_EMU_TESTS=1 in the environment of the pytest it spawns)."""
    if not any(it.get_closest_marker("gpu") for it in items):
        return
    lib_path = os.environ.get("B2R_LIBRARY")
    if not lib_path or os.environ.get("B2R_EMU_TESTS") == "1":
        return
    import ctypes
    try:
        emulated = hasattr(ctypes.CDLL(lib_path), "emu_enqueue_kernel")
    except OSError:
        emulated = False
    if emulated:
        raise pytest.UsageError(f"-m gpu tests refuse B2R_LIBRARY={lib_path}: it is the CPU emulation build (tools/emu), not the product on a GPU; "
                                "tests/test_cpu_emu.py runs these files on it deliberately (B2R_EMU_TESTS=1)")
