"""Shared fixtures.  `-m "not gpu"` runs the oracle against the reference's analytic known answers,
the host logic, the C-ABI symbol check and the kernel sources under the CPU-thread emulation
(tests/emu, test infrastructure only).  `-m gpu` runs the parity tests proper through the C ABI of
the nvcc-built library on a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TABLES = os.path.join(ROOT, "tests", "golden", "tables")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def tables():
    return TABLES


@pytest.fixture(scope="session")
def emu_lib():
    """The kernel sources compiled for CPU threads (tests only; never loaded by the product)."""
    from specter_b200 import api, build
    # SPECTER_EMU_LIB: another build of the same sources, e.g. the AddressSanitizer one of tools/emu_asan.sh
    return api.Library(os.environ.get("SPECTER_EMU_LIB") or build.build_emu())


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from specter_b200 import api
    return api.load_library()


def rel_err(x, y):
    d = float(np.abs(np.asarray(y)).max())
    return float(np.abs(np.asarray(x) - np.asarray(y)).max()) / (d if d > 0 else 1.0)
