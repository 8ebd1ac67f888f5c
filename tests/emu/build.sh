#!/bin/bash
# Builds tests/emu/libtde_emu.so: the product's CUDA sources compiled for the host-side lockstep emulator
# (cuda_emu.h).  Test infrastructure only - see the header of cuda_emu.h.
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
g++ -x c++ -std=c++17 -O2 -g -fPIC -shared -DTDE_HOST_EMU -ffp-contract=off -fno-fast-math -mfma \
    -Wl,-Bsymbolic -Wno-unknown-pragmas -Wno-attributes -I"$here" "$@" -o "$here/libtde_emu.so" "$root/torchdriveenv_b200/csrc/tde_b200.cu"
echo "built $here/libtde_emu.so"
