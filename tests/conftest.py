import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# MRTS_EMU=1 runs the `gpu` tests against tests/emu (coroutine-emulated warps) for debugging without a GPU.
EMU = os.environ.get("MRTS_EMU", "0") == "1"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if EMU:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import emu_backend
        emu_backend.use_emulator()


@pytest.fixture(scope="session")
def backend():
    """'cuda' or 'emu'.  With the CUDA backend the tests fail loudly if the library or the device is missing."""
    if EMU:
        return "emu"
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device (or MRTS_EMU=1 for the emulator)"
    from microrts_b200 import _ffi
    _ffi.lib()
    return "cuda"


@pytest.fixture(scope="session")
def maps():
    import golden_io
    return golden_io.load_maps()


@pytest.fixture(scope="session")
def traces():
    import golden_io
    return golden_io.load_traces()
