"""TEST TOOLING ONLY: point the microrts_b200 ctypes binding at the coroutine-emulated build of the engine
(tests/emu/libmicrorts_emu.so) so the parity tests can be debugged on a machine without a GPU."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def use_emulator(rebuild=True):
    so = os.path.join(HERE, "libmicrorts_emu.so")
    if os.environ.get("MRTS_EMU_ASAN", "0") == "1":
        # AddressSanitizer + UBSan build of the same sources (build.sh asan); run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
        so = os.path.join(HERE, "libmicrorts_emu_asan.so")
        if rebuild:
            subprocess.check_call([os.path.join(HERE, "build.sh"), "asan"])
    elif rebuild:
        subprocess.check_call([os.path.join(HERE, "build.sh")])
    from microrts_b200 import _ffi
    _ffi._lib = _ffi.bind(ctypes.CDLL(so))
    return _ffi._lib
