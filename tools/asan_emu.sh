#!/bin/bash
# Host-compiled kernel bodies + orchestration under AddressSanitizer: the CPU test files that drive them, with the sanitizer build in place of
# tests/emu's.  Found the stale-queue read after an arena overflow (round 2).  Usage: bash tools/asan_emu.sh [pytest args]
set -e
cd "$(dirname "$0")/.."
S=sailor_b200/csrc
g++ -std=c++17 -O2 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared -DSPT_EMU -I include -o /tmp/libsailor_pt_emu_asan.so \
    -x c++ $S/capi.cu -x c++ $S/backend.cu $S/gltf_loader.cpp $S/png_codec.cpp $S/jpeg_codec.cpp $S/image_io.cpp -lz
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 SAILOR_EMU_LIB=/tmp/libsailor_pt_emu_asan.so \
    python -m pytest tests/test_host_logic.py tests/test_multi_device.py tests/test_gltf_containers.py -x -q -m "not gpu" "$@"
