#!/bin/bash
# Replays every CPU emulation test with tests/emu built under AddressSanitizer: shared memory, TMA boxes and global arrays are
# exact-size host allocations there, so an out-of-bounds index in a kernel body (the code the CUDA kernels execute) is reported.
#   bash tools/asan_emu.sh        (about 2 minutes; restores the regular build afterwards)
set -eu
cd "$(dirname "$0")/.."
so=tests/emu/libpxb_emu.so
cp "$so" /tmp/libpxb_emu.so.keep
trap 'cp /tmp/libpxb_emu.so.keep "$so"' EXIT
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -fPIC -shared -x c++ -o "$so" tests/emu/pxb_emu.cpp
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
    python -m pytest tests/test_emu_iter.py tests/test_emu_kernels.py tests/test_emu_stencil_tiled.py tests/test_slab_cpu.py \
    tests/test_emu_device_solvers.py -x -q -p no:cacheprovider
# (tests/test_slab_cpu.py::test_slab_worker_on_the_emulated_device launches torchrun: the children inherit LD_PRELOAD / ASAN_OPTIONS)
