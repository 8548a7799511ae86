#!/bin/sh
# compute-sanitizer is closed on this GPU pool, so out-of-bounds / UB checks of the DEVICE code run on
# the CPU: the same oc_device.cuh compiled for the host (tests/emu) with ASan + UBSan, driven through
# every golden trace plus auto-reset / terminal-obs / rollout / ragged-size cases.
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -shared -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer \
    -Itests/emu tests/emu/oc_emu.cpp -o /tmp/liboc_emu_asan.so
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
    python tools/asan_emu_run.py /tmp/liboc_emu_asan.so
