#!/bin/sh
# TEST TOOLING ONLY: compiles the device engine for the host with coroutine-emulated warps (tests/emu/cuda_shim.h).
# The result (tests/emu/libmicrorts_emu.so) is used to debug warp-synchronous logic without a GPU; it is never part of
# the product and microrts_b200 cannot load it.
set -e
cd "$(dirname "$0")/../.."
if [ "$1" = "asan" ]; then
    # the same sources under AddressSanitizer + UndefinedBehaviorSanitizer: out-of-bounds accesses of the emulated shared-memory
    # window and of every "device" buffer, signed overflow, bad shifts (tools/sanitize.py is the workload)
    [ tests/emu/libmicrorts_emu_asan.so -nt microrts_b200/csrc/engine.cuh ] && [ tests/emu/libmicrorts_emu_asan.so -nt microrts_b200/csrc/scripted.cuh ] && \
        [ tests/emu/libmicrorts_emu_asan.so -nt microrts_b200/csrc/microrts_cuda.cu ] && exit 0
    exec g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -DMRTS_EMU -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unused-function -Wno-unknown-pragmas \
        -I tests/emu -I microrts_b200/csrc -x c++ microrts_b200/csrc/microrts_cuda.cu -o tests/emu/libmicrorts_emu_asan.so
fi
g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -DMRTS_EMU -Wall -Wno-unused-function -Wno-unknown-pragmas \
    -I tests/emu -I microrts_b200/csrc -x c++ microrts_b200/csrc/microrts_cuda.cu -o tests/emu/libmicrorts_emu.so
